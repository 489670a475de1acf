"""Seeded scorer weights (no checkpoints are reachable offline).

mlp2x512 (BASELINE configs 2-5): W ~ N(0, 1/fan_in) ("variance_scaling fan_in normal",
NANN_impls/nann/model/model_util.py:48), biases 0.1 (:49).

Attention scorer (config 1) as ONE fp32 blob in this order (TF kernel layout [in][out]):
  Wq1[64,128] bq1[128] alpha_q[128]  Wq2[128,256] bq2[256]      nonlinear_attention dense, dense_1, prelu_q
  Wk1[64,128] bk1[128] alpha_k[128]  Wk2[128,256] bk2[256]      dense_2, dense_3, prelu_k
  1_dnn: W[128,128] b[128] bn_scale[128] bn_shift[128] alpha[128]
  2_dnn: W[128,64]  b[64]  bn_scale[64]  bn_shift[64]  alpha[64]
  3_dnn: W[64,32]   b[32]  bn_scale[32]  bn_shift[32]  alpha[32]
  4_dnn: W[32]                                                   (no bias, model.py:220)
BatchNorm (tf.layers.batch_normalization, inference, eps=1e-3) is folded here:
  bn_scale = gamma / sqrt(var + 1e-3), bn_shift = beta - mean * bn_scale   (float32).
"""
import numpy as np

ATT_BLOB = (64 * 128 + 128 + 128 + 128 * 256 + 256) * 2 + 128 * 128 + 4 * 128 + 128 * 64 + 4 * 64 + 64 * 32 + 4 * 32 + 32


def mlp_weights(d=128, H=512, seed=3):
    rng = np.random.default_rng(seed)
    W1 = (rng.standard_normal((H, 2 * d)) / np.sqrt(2 * d)).astype(np.float32)
    W2 = (rng.standard_normal((H, H)) / np.sqrt(H)).astype(np.float32)
    w3 = (rng.standard_normal(H) / np.sqrt(H)).astype(np.float32)
    b1 = np.full(H, 0.1, np.float32)
    b2 = np.full(H, 0.1, np.float32)
    return W1, b1, W2, b2, w3


def fold_bn(gamma, beta, mean, var, eps=1e-3):
    scale = (gamma.astype(np.float32) / np.sqrt(var.astype(np.float32) + np.float32(eps))).astype(np.float32)
    shift = (beta.astype(np.float32) - mean.astype(np.float32) * scale).astype(np.float32)
    return scale, shift


def attention_blob(seed=3):
    rng = np.random.default_rng(seed)
    parts = []

    def dense(i, o, bias_init=None):
        parts.append((rng.standard_normal((i, o)) / np.sqrt(i)).astype(np.float32).ravel())
        if bias_init is not None:
            parts.append((bias_init + 0.05 * rng.standard_normal(o)).astype(np.float32))

    def alpha(n):
        parts.append((0.25 + 0.05 * rng.standard_normal(n)).astype(np.float32))

    for _ in range(2):  # q side, then k side
        dense(64, 128, 0.0); alpha(128); dense(128, 256, 0.0)
    for i, o in ((128, 128), (128, 64), (64, 32)):
        dense(i, o, 0.1)
        g = rng.uniform(0.5, 1.5, o); b = 0.1 * rng.standard_normal(o)
        m = 0.1 * rng.standard_normal(o); v = rng.uniform(0.5, 1.5, o)
        sc, sh = fold_bn(g, b, m, v)
        parts.append(sc); parts.append(sh)
        alpha(o)
    parts.append((rng.standard_normal(32) / np.sqrt(32)).astype(np.float32))
    blob = np.concatenate(parts).astype(np.float32)
    assert blob.size == ATT_BLOB
    return blob
