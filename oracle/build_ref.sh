#!/bin/bash
# TEST INFRASTRUCTURE: compiles the reference's own Cython batch packer (fairseq/data/data_utils_fast.pyx, the native
# code behind fairseq.data.data_utils.batch_by_size) from the sources where they lie under /root/reference into
# oracle/_ref/ (git-ignored).  Used only to pin espresso_b200's batch packer (oracle/pin_against_reference.py, section
# "batching"); nothing under espresso_b200/ loads it.  No reference source is copied into the repository.
set -e
REF=${1:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
mkdir -p "$OUT"
PYINC=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NPINC=$(python -c "import numpy; print(numpy.get_include())")
EXT=$(python -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
python -m cython --cplus -3 -o "$OUT/data_utils_fast.cpp" "$REF/fairseq/data/data_utils_fast.pyx"
g++ -O2 -shared -fPIC -std=c++14 -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION -I"$PYINC" -I"$NPINC" \
    "$OUT/data_utils_fast.cpp" -o "$OUT/data_utils_fast$EXT"
rm -f "$OUT/data_utils_fast.cpp"
echo "built $OUT/data_utils_fast$EXT"
