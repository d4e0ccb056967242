"""CPU restatement of the value network of code/utils/alpha_nnet.py:19-56 (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED against TensorFlow: TF/Keras cannot be installed in the build container and the reference ships no
weights or known answers, so this restatement follows the Keras 2.x layer definitions:
NHWC input, Conv2D kernels HWIO with 'same' zero padding and no bias, BatchNormalization(axis=3) in inference mode
y = gamma * (x - mean) / sqrt(var + 1e-3) + beta, residual add after the second BN and before the ReLU, 1x1 conv to
one channel + BN + ReLU, Flatten in (h, w) order, Dense(128) + bias + ReLU, Dense(3) + bias + tanh.
What CAN be pinned, and is, is the CUDA network against this restatement (float64 accumulation).
"""
import numpy as np

K = 128
BN_EPS = 1e-3


def init_weights(side, seed=0, dtype=np.float32, randomize_bn=False):
    """Glorot-uniform kernels, zero biases, BN at init (Keras defaults).  randomize_bn perturbs the BN statistics so
    that tests exercise the folding arithmetic."""
    rng = np.random.default_rng(seed)
    n = 2 * side - 1

    def glorot(shape, fan_in, fan_out):
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(dtype)

    def bn(c):
        if randomize_bn:
            return dict(gamma=rng.uniform(0.5, 1.5, c).astype(dtype), beta=rng.uniform(-0.3, 0.3, c).astype(dtype),
                        mean=rng.uniform(-0.2, 0.2, c).astype(dtype), var=rng.uniform(0.5, 1.5, c).astype(dtype))
        return dict(gamma=np.ones(c, dtype), beta=np.zeros(c, dtype), mean=np.zeros(c, dtype), var=np.ones(c, dtype))

    w = {"side": side}
    w["conv0"] = glorot((3, 3, 3, K), 3 * 9, K * 9)                      # alpha_nnet.py:21
    w["bn0"] = bn(K)
    for b in range(4):                                                    # alpha_nnet.py:24-47
        for j in range(2):
            w["res%d_conv%d" % (b, j)] = glorot((3, 3, K, K), K * 9, K * 9)
            w["res%d_bn%d" % (b, j)] = bn(K)
    w["head_conv"] = glorot((1, 1, K, 1), K, 1)                           # alpha_nnet.py:49
    w["head_bn"] = bn(1)
    w["dense1_w"] = glorot((n * n, K), n * n, K)                          # alpha_nnet.py:52
    w["dense1_b"] = (rng.uniform(-0.1, 0.1, K).astype(dtype) if randomize_bn else np.zeros(K, dtype))
    w["dense2_w"] = glorot((K, 3), K, 3)                                  # alpha_nnet.py:54
    w["dense2_b"] = (rng.uniform(-0.1, 0.1, 3).astype(dtype) if randomize_bn else np.zeros(3, dtype))
    return w


def _conv_same(x, k):
    """x [B,H,W,Cin] float64, k [kh,kw,Cin,Cout] -> [B,H,W,Cout] (stride 1, zero 'same' padding)."""
    kh, kw, cin, cout = k.shape
    B, H, W, _ = x.shape
    ph, pw = kh // 2, kw // 2
    xp = np.zeros((B, H + 2 * ph, W + 2 * pw, cin), x.dtype)
    xp[:, ph:ph + H, pw:pw + W] = x
    out = np.zeros((B, H, W, cout), x.dtype)
    for dy in range(kh):
        for dx in range(kw):
            out += xp[:, dy:dy + H, dx:dx + W].reshape(-1, cin).dot(k[dy, dx].astype(x.dtype)).reshape(B, H, W, cout)
    return out


def _bn(x, p):
    g, b, m, v = (p[k].astype(x.dtype) for k in ("gamma", "beta", "mean", "var"))
    return g * (x - m) / np.sqrt(v + BN_EPS) + b


def bf16_round(a):
    """round-to-nearest-even to bfloat16, returned as float64 (what the bf16 tensor-core path stores)."""
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(np.float64).reshape(f.shape)


def forward_layers(w, X, bf16=True):
    """outputs of the 9 tower convolutions (after BN / residual / ReLU; the last one after the 1x1 head conv too)."""
    x = np.asarray(X, dtype=np.float64)
    q = bf16_round if bf16 else (lambda a: a)
    relu = lambda a: np.maximum(a, 0)
    outs = []
    h = q(relu(_bn(_conv_same(q(x), q(w["conv0"])), w["bn0"])))
    outs.append(h)
    for b in range(4):
        sc = h
        h = q(relu(_bn(_conv_same(h, q(w["res%d_conv0" % b])), w["res%d_bn0" % b])))
        outs.append(h)
        h = q(relu(_bn(_conv_same(h, q(w["res%d_conv1" % b])), w["res%d_bn1" % b]) + sc))
        outs.append(h)
    outs[-1] = relu(_bn(_conv_same(h, w["head_conv"]), w["head_bn"]))
    return outs


def forward(w, X, dtype=np.float64, bf16=False):
    """X [B, n, n, 3] -> [B, 3] raw tanh outputs (no obstacle mask).
    bf16=True emulates the storage precision of the CUDA path: bf16 input planes, conv kernels and layer outputs
    (exact products, wide accumulation), fp32/64 everywhere else."""
    x = np.asarray(X, dtype=dtype)
    q = bf16_round if bf16 else (lambda a: a)
    relu = lambda a: np.maximum(a, 0)
    x = q(x)
    h = q(relu(_bn(_conv_same(x, q(w["conv0"])), w["bn0"])))
    for b in range(4):
        sc = h
        h = q(relu(_bn(_conv_same(h, q(w["res%d_conv0" % b])), w["res%d_bn0" % b])))
        h = q(relu(_bn(_conv_same(h, q(w["res%d_conv1" % b])), w["res%d_bn1" % b]) + sc))
    h = relu(_bn(_conv_same(h, w["head_conv"]), w["head_bn"]))
    h = h.reshape(h.shape[0], -1)
    h = relu(h.dot(w["dense1_w"].astype(dtype)) + w["dense1_b"].astype(dtype))
    return np.tanh(h.dot(w["dense2_w"].astype(dtype)) + w["dense2_b"].astype(dtype))


def v(w, X, numpy1_mask=False):
    """AlphaNNet.v (alpha_nnet.py:61-76): forward + obstacle mask."""
    X = np.asarray(X, dtype=np.float32)
    V = forward(w, X).astype(np.float32)
    cy, cx = X.shape[1] // 2, X.shape[2] // 2
    thr = 0.04 if numpy1_mask else np.float32(0.04)
    V[X[:, cy, cx - 1, 1] >= thr, 0] = -1.0
    V[X[:, cy - 1, cx, 1] >= thr, 1] = -1.0
    V[X[:, cy, cx + 1, 1] >= thr, 2] = -1.0
    return V


FLOPS_PER_EVAL = {11: 1043724288, 19: 3240040448}   # 2*MAC, convs + dense


def flops_per_eval(side):
    n = 2 * side - 1
    macs = n * n * (9 * 3 * K + 8 * 9 * K * K + K) + n * n * K + K * 3
    return 2 * macs
