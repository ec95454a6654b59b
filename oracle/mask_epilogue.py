"""oracle/mask_epilogue.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement (float32 step by step, like the ATen kernels) of the tail of RMNet.segment / RMNet.forward after the
decoder, batch 1:
    models/rmnet.py:368-370  ps = F.softmax(logits, dim=1)[:, 1]
    models/rmnet.py:289-302  soft_aggregation (bg = prod(1 - ps), clamp(1e-7, 1 - 1e-7), log(em / (1 - em)))
    models/rmnet.py:376-380  un-pad
    models/rmnet.py:436-448  new-object / non-existing-object channel overrides
    models/rmnet.py:450      est_masks[:, t] = F.softmax(logit, dim=1)
Pinned by tests/golden/mask_epilogue.npz (produced by the reference's own soft_aggregation + torch ops).
"""
import numpy as np

CH_KEEP, CH_ABSENT, CH_NEW = 0, 1, 2
F32 = np.float32


def mask_epilogue(dec_logits, K, frame_hw, channel_modes=None, new_mask=None, dtype=np.float32):
    """dec_logits [n,2,Hp,Wp] -> (logit [1,K,H,W], est_mask [1,K,H,W]); dtype=np.float64 gives the exact-arithmetic answer."""
    x = np.asarray(dec_logits, dtype)
    n, _, Hp, Wp = x.shape
    H, W = frame_hw
    lh, lw = (Hp - H) // 2, (Wp - W) // 2                      # utils/helpers.py:105-124 (lower pads are the floor halves)
    m = x.max(1, keepdims=True)                                 # :368
    e = np.exp(x - m)
    ps = (e / e.sum(1, keepdims=True))[:, 1]                    # :370
    em = np.zeros((K, Hp, Wp), dtype)                           # :293
    em[0] = np.prod(1 - ps, axis=0)                             # :297
    em[1:n + 1] = ps                                            # :298
    # :300 -- torch.clamp casts its python-float bounds to the tensor's float32: the upper bound is 1 - 2^-23, which
    # is what makes max(logit) = 15.9424 (the reference's own comment at :441), not 16.1181
    em = np.clip(em, dtype(F32(1e-7)), dtype(F32(1 - 1e-7)))
    logit = np.log(em / (1 - em))                               # :301
    logit = logit[:, lh:lh + H, lw:lw + W].copy()               # :376-380
    for j in range(K):
        mode = CH_KEEP if channel_modes is None else channel_modes[j]
        if mode == CH_NEW:
            logit[j] = np.asarray(new_mask[j], dtype) * dtype(32.0605) - dtype(16.1181)   # :442
        elif mode == CH_ABSENT:
            logit[j] = dtype(-16.1181)                                                      # :448
    z = np.exp(logit - logit.max(0, keepdims=True))             # :450
    est = z / z.sum(0, keepdims=True)
    return logit[None].astype(np.float32), est[None].astype(np.float32)
