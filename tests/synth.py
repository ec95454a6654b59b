"""Seeded synthetic inputs shared by the golden generator, the CPU tests and the GPU parity tests.

Everything is derived from numpy's PCG64 ``default_rng(seed)`` (stream-stable by numpy policy), so
fixtures only need to store seeds, shapes and the reference's OUTPUTS.
"""
import numpy as np

CK, CV = 128, 512  # models/rmnet.py:185-186


def rect_label_map(rng, n_obj, H, W, lo=0.15, hi=0.45):
    """SURVEY 8d synthetic clip frame: n axis-aligned rectangles painted in order o = 1..n."""
    lab = np.zeros((H, W), np.int64)
    for o in range(1, n_obj + 1):
        bh = max(1, int(rng.uniform(lo, hi) * H))
        bw = max(1, int(rng.uniform(lo, hi) * W))
        y0 = int(rng.integers(0, H - bh + 1))
        x0 = int(rng.integers(0, W - bw + 1))
        lab[y0:y0 + bh, x0:x0 + bw] = o
    return lab


def onehot(lab, K):
    return np.stack([(lab == k) for k in range(K)]).astype(np.float32)


def soft_masks(rng, lab, K, sharp=8.0):
    """Softmax-like probabilities around a label map (est_masks of later frames are soft)."""
    logits = rng.standard_normal((K,) + lab.shape).astype(np.float32)
    logits += sharp * onehot(lab, K)
    e = np.exp(logits - logits.max(0, keepdims=True))
    return (e / e.sum(0, keepdims=True)).astype(np.float32)


def flow_field(rng, H, W, sigma=2.0, half_pixel=False):
    f = (rng.standard_normal((2, H, W)) * sigma).astype(np.float32)
    if half_pixel:
        f = (np.round(f * 2) / 2).astype(np.float32)
    return f


def memory_read_inputs(seed, n, T, h, w, scale=1.0):
    rng = np.random.default_rng(seed)
    m_key = (rng.standard_normal((n, CK, T, h, w)) * scale).astype(np.float32)
    m_val = rng.standard_normal((n, CV, T, h, w)).astype(np.float32)
    q_key = (rng.standard_normal((n, CK, h, w)) * scale).astype(np.float32)
    q_val = rng.standard_normal((n, CV, h, w)).astype(np.float32)
    return m_key, m_val, q_key, q_val


def cell_boxes(rng, n, T, h, w, frac=(0.3, 0.8)):
    """Random low-res cell rectangles [n,T,4] = (cx0, cx1, cy0, cy1) inclusive."""
    out = np.zeros((n, T, 4), np.int32)
    for o in range(n):
        for t in range(T):
            bw = max(1, int(rng.uniform(*frac) * w))
            bh = max(1, int(rng.uniform(*frac) * h))
            x0 = int(rng.integers(0, w - bw + 1))
            y0 = int(rng.integers(0, h - bh + 1))
            out[o, t] = (x0, x0 + bw - 1, y0, y0 + bh - 1)
    return out


def affine_pair(rng):
    """Two plausible RandomAffine matrices (utils/data_transforms.py:262-302 style): small rotation/scale/shift."""
    ms = []
    for _ in range(2):
        a = rng.uniform(-0.3, 0.3)
        s = rng.uniform(0.8, 1.25)
        m = np.array([[s * np.cos(a), -s * np.sin(a), rng.uniform(-20, 20)],
                      [s * np.sin(a), s * np.cos(a), rng.uniform(-20, 20)]], np.float32)
        ms.append(m)
    return ms


def decoder_logits(rng, n_obj, H, W, sharp=6.0):
    """Decoder output [n,2,Hp,Wp] on the padded frame: object o is likely inside its rectangle, noise elsewhere;
    a band of exact ties and a band of saturated pixels exercise the clamp (models/rmnet.py:300)."""
    Hp, Wp = (H + 15) // 16 * 16, (W + 15) // 16 * 16
    x = rng.standard_normal((n_obj, 2, Hp, Wp)).astype(np.float32)
    lab = rect_label_map(rng, n_obj, Hp, Wp)
    for o in range(n_obj):
        x[o, 1] += sharp * (lab == o + 1)
        x[o, 0] += sharp * (lab != o + 1)
    x[:, :, :2, :] = 0.0                 # ties: ps = 0.5 exactly
    x[:, 1, 2:4, :] += 40.0              # saturated foreground: ps -> 1, clamped
    x[:, 0, 4:6, :] += 40.0              # saturated background: ps -> 0, clamped
    return x


def epilogue_logit_tolerance(x, K, H, W, base=1e-3, c=8.0):
    """Per-element bound for comparing two float32 evaluations of models/rmnet.py:289-302 on decoder logits x [n,2,Hp,Wp].
    log(em / (1 - em)) cancels catastrophically where an object is saturated: one float32 ulp of ps moves 1 - ps by a
    relative 2^-24 / (1 - ps), and the background product collects that from every object.  Where no object is
    saturated the bound is `base` (north_star's 1e-3); elsewhere it grows with the amplification that ANY float32
    implementation of the reference's formula is subject to (torch CPU vs torch CUDA differ by as much)."""
    x = np.asarray(x, np.float64)
    n, _, Hp, Wp = x.shape
    lh, lw = (Hp - H) // 2, (Wp - W) // 2
    ps = 1.0 / (1.0 + np.exp(x[:, 0] - x[:, 1]))
    amp = 2.0 ** -24 / np.maximum(1.0 - ps, 2.0 ** -24)
    tol = np.full((K, Hp, Wp), base)
    tol[0] += c * amp.sum(0)
    tol[1:n + 1] += c * amp
    return tol[None, :, lh:lh + H, lw:lw + W]
