"""Minimal stand-in for omegaconf: just enough surface for `import fairseq` from
/root/reference in the authoring container (no network, omegaconf not installed).
Test infrastructure only -- never imported by the product path."""
__version__ = "2.0.6"
MISSING = "???"


def II(s):
    return "${" + s + "}"


class _Node(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class DictConfig(_Node):
    pass


class ListConfig(list):
    pass


class _Ctx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def open_dict(cfg):
    return _Ctx()


def read_write(cfg):
    return _Ctx()


class OmegaConf:
    @staticmethod
    def create(obj=None, **kw):
        return obj

    @staticmethod
    def is_config(obj):
        return isinstance(obj, (DictConfig, ListConfig))

    @staticmethod
    def set_struct(cfg, flag):
        pass

    @staticmethod
    def to_container(cfg, resolve=False, **kw):
        return cfg

    @staticmethod
    def merge(*cfgs):
        """Later configs override earlier ones.  The shim only ever merges a default-constructed dataclass with a fully
        populated config of the same class (fairseq/dataclass/utils.py merge_with_parent), so "the last one wins"."""
        import copy

        out = [c for c in cfgs if c is not None][-1]
        return copy.copy(out)

    @staticmethod
    def structured(obj):
        return obj

    @staticmethod
    def is_dict(obj):
        return isinstance(obj, dict)

    @staticmethod
    def is_list(obj):
        return isinstance(obj, list)
