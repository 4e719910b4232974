"""Definition of the full-size parity fixture (tests/golden/fullsize_conformer.npz): model configuration, seeded inputs
and the list of gradients stored.  TEST INFRASTRUCTURE ONLY -- shared by the generator (oracle/pin_against_reference.py,
which runs the REAL reference) and by the CPU / GPU parity tests; it needs neither the reference tree nor a GPU."""
import numpy as np

FULLSIZE_CFG = dict(embed_dim=512, ffn_dim=2048, heads=8, layers=17, layer_type="conformer", dw_kernel=31, dropout=0.0,
                    attention_dropout=0.0, activation_dropout=0.0, layernorm_embedding=True, final_layer_norm=False, vocab=5004)
FULLSIZE_GRADS = ["encoder.fc_out.bias", "encoder.fc0.bias", "encoder.layernorm_embedding.weight",
                  "encoder.layers.0.self_attn.pos_bias_u", "encoder.layers.8.self_attn.pos_bias_v",
                  "encoder.layers.16.self_attn.out_proj.bias", "encoder.layers.0.ffn1.w_1.bias",
                  "encoder.layers.16.ffn2.w_2.bias", "encoder.layers.8.conv_module.depthwise_conv.weight",
                  "encoder.layers.8.conv_module.batch_norm.weight", "encoder.layers.16.final_layer_norm.weight",
                  "encoder.pre_encoder.batchnorms.3.weight"]
FULLSIZE_GRADS_SUB = ["encoder.layers.0.self_attn.q_proj.weight", "encoder.layers.16.self_attn.v_proj.weight",
                      "encoder.layers.8.self_attn.pos_proj.weight", "encoder.layers.8.ffn1.w_1.weight",
                      "encoder.layers.16.ffn2.w_2.weight", "encoder.layers.0.conv_module.pointwise_conv1.weight",
                      "encoder.fc_out.weight", "encoder.fc0.weight", "encoder.pre_encoder.convolutions.2.weight"]


def fullsize_inputs():
    """Seeded inputs of the full-size fixture (shared by the generator and the tests): 3 utterances of 3.1 / 6.4 / 10 s."""
    rs = np.random.RandomState(2024)
    lens = np.array([998, 638, 308], dtype=np.int64)
    feats = rs.randn(3, 998, 80).astype(np.float16).astype(np.float32)
    for b in range(3):
        feats[b, lens[b]:] = 0.0
    tgt = np.full((3, 42), 1, dtype=np.int64)
    for b, u in enumerate((40, 26, 12)):
        tgt[b, :u] = rs.randint(4, 5004, size=u)
        tgt[b, u] = 2
    return feats, lens, tgt


def fullsize_cotangent(out_lens, V=5004):
    """Fixed pseudo-random cotangent G [B, T', V] (zero on padded frames).  Gradients are compared for the LINEAR
    functional sum(G * logits): d/dlogits = G exactly, so the comparison measures the encoder's backward pass and not the
    chaotic sensitivity of CTC alignment posteriors of a randomly initialised model (where a 5 % bf16 logit perturbation
    moves 10-50 % of the occupancy between neighbouring labels -- in the reference's own bf16 run too)."""
    rs = np.random.RandomState(77)
    B, T = len(out_lens), int(max(out_lens))
    G = (rs.standard_normal((B, T, V)) * 0.05).astype(np.float32)
    for b, l in enumerate(out_lens):
        G[b, int(l):] = 0.0
    return G
