"""oracle/memory_read.py -- TEST INFRASTRUCTURE ONLY.

numpy (BLAS-backed) restatement of the reference's floating-point memory read and of the
regional masking that surrounds it.  Used as the parity checker and as bench.py's CPU
baseline; never imported by the product.
"""
import math

import numpy as np


def memory_read(m_key, m_val, q_key, q_val, dtype=np.float32, want_p=False):
    """MemoryReader.forward, models/rmnet.py:147-165 -- same op order as the reference:
    bmm(K^T, Q) (:155) -> / sqrt(D_e) (:156) -> softmax over dim=1 (:157) -> bmm(V, p) (:160)
    -> cat([mem, q_val], 1) (:163).  dtype=np.float32 mirrors the reference arithmetic;
    dtype=np.float64 is the floor used to separate our error from the reference's own.
      m_key [n,Ck,T,h,w] m_val [n,Cv,T,h,w] q_key [n,Ck,h,w] q_val [n,Cv,h,w]
      -> mem_val [n,2*Cv,h,w] (f32), p [n,T*h*w,h*w] | None
    """
    n, Ck, T, h, w = m_key.shape
    Cv = m_val.shape[1]
    M, N = T * h * w, h * w
    out = np.empty((n, Cv + q_val.shape[1], h, w), np.float32)
    ps = [] if want_p else None
    for o in range(n):  # one object at a time: p is [M,N] (2 GB / object at 720p, T=40)
        mi = np.ascontiguousarray(m_key[o].reshape(Ck, M).T, dtype=dtype)  # :151-152
        qi = np.asarray(q_key[o].reshape(Ck, N), dtype=dtype)              # :153
        p = mi @ qi                                                        # :155
        p /= dtype(math.sqrt(Ck))                                          # :156
        p -= p.max(axis=0, keepdims=True)                                  # :157 softmax(dim=1)
        np.exp(p, out=p)
        p /= p.sum(axis=0, keepdims=True, dtype=dtype)
        mo = np.asarray(m_val[o].reshape(Cv, M), dtype=dtype)              # :158
        mem = mo @ p                                                       # :160
        out[o, :Cv] = mem.reshape(Cv, h, w)
        out[o, Cv:] = q_val[o]                                             # :163
        if want_p:
            ps.append(p.astype(np.float32))
    return out, (np.stack(ps) if want_p else None)


def regional_mask_memory(k4, v4, att_map_padded):
    """models/rmnet.py:245-248: att16 = interpolate(att_map, 1/16); k4 *= att16; v4 *= att16.
      k4 [B,K,Ck,1,h,w], v4 [B,K,Cv,1,h,w], att_map_padded [B,K,Hp,Wp] -> masked (k4, v4)"""
    from . import downsample16
    att16 = downsample16(att_map_padded)[:, :, None, None]   # unsqueeze(2).unsqueeze(2)
    return k4 * att16, v4 * att16


def regional_mask_query(k4e, v4e, att_map_padded_obj):
    """models/rmnet.py:355-358.  k4e [n,Ck,h,w], v4e [n,Cv,h,w], att_map_padded_obj [n,1,Hp,Wp]."""
    from . import downsample16
    att16 = downsample16(att_map_padded_obj)
    return k4e * att16, v4e * att16


def regional_memory_read(m_key_raw, m_val_raw, mem_att_padded, q_key_raw, q_val_raw, q_att_padded,
                         dtype=np.float32):
    """The regional read exactly as the reference composes it (models/rmnet.py:243-248 per memory
    frame, :355-361 per query frame): mask, then the DENSE reader.
      m_key_raw [n,Ck,T,h,w], m_val_raw [n,Cv,T,h,w]  unmasked per-object memory K/V
      mem_att_padded [n,T,Hp,Wp]   full-res att map of each (object, memory frame), padded coordinates
      q_key_raw [Ck,h,w], q_val_raw [Cv,h,w]           one query frame (expanded to n objects, :332-333)
      q_att_padded [n,Hp,Wp]       full-res att map of each object on the query frame (already padded, :307)
    """
    from . import downsample16
    n = m_key_raw.shape[0]
    a_m = downsample16(mem_att_padded)                      # [n,T,h,w]
    mk = m_key_raw * a_m[:, None]
    mv = m_val_raw * a_m[:, None]
    a_q = downsample16(q_att_padded)                        # [n,h,w]
    qk = np.broadcast_to(q_key_raw, (n,) + q_key_raw.shape) * a_q[:, None]
    qv = np.broadcast_to(q_val_raw, (n,) + q_val_raw.shape) * a_q[:, None]
    return memory_read(mk.astype(np.float32), mv.astype(np.float32), qk.astype(np.float32),
                       qv.astype(np.float32), dtype=dtype)[0]
