"""TEST INFRASTRUCTURE: the REFERENCE's composition of one frame of the regional memory-read path, run on the GPU with
torch's own CUDA ops and the reference's unmodified CUDA extension (oracle/_ref, compiled from /root/reference by
oracle/Makefile).  Restates models/rmnet.py line by line (the model file itself does not travel to the GPU box):

    memorize side : pad (:212) -> get_att_map (:244, reference extension) -> interpolate 1/16 (:245) -> k4*att, v4*att (:247-248)
                    -> pad_memory (:191-205) -> cat to the bank (:420-421)
    segment side  : warp (:252-278) -> get_att_map (:286, reference extension) -> pad (:307) -> interpolate (:356)
                    -> k4e*att, v4e*att (:357-358) -> MemoryReader.forward (:147-165)
Used by tests/test_gpu_parity.py (full-size parity of RegionalMemory.step) and tools/ref_gpu_step.py (GPU-vs-GPU timing).
"""
import math

import torch
import torch.nn.functional as F


def pad16(x, H, W):
    Hp, Wp = (H + 15) // 16 * 16, (W + 15) // 16 * 16
    lh, lw = (Hp - H) // 2, (Wp - W) // 2
    return F.pad(x, (lw, Wp - W - lw, lh, Hp - H - lh))


def torch_warp(img0, flow):
    """RMNet.warp restated with the same torch ops on the GPU (models/rmnet.py:252-278)."""
    B, C, H, W = img0.size()
    x_axis = torch.arange(0, W).view(1, -1).repeat(H, 1).view(1, 1, H, W).repeat(B, 1, 1, 1)
    y_axis = torch.arange(0, H).view(-1, 1).repeat(1, W).view(1, 1, H, W).repeat(B, 1, 1, 1)
    grid = torch.cat((x_axis, y_axis), 1).float().to(img0.device)
    vgrid = grid + flow
    vgrid[:, 0, :, :] = 2.0 * vgrid[:, 0, :, :].clone() / max(W - 1, 1) - 1.0
    vgrid[:, 1, :, :] = 2.0 * vgrid[:, 1, :, :].clone() / max(H - 1, 1) - 1.0
    vgrid = vgrid.permute(0, 2, 3, 1)
    img1 = F.grid_sample(img0.clone(), vgrid, align_corners=True)
    mask = F.grid_sample(torch.ones_like(img0), vgrid, align_corners=True)
    mask[mask < 0.9999] = 0
    mask[mask > 0] = 1
    return img1 * mask, mask


class ReferenceClip:
    """Memory bank + per-frame step exactly as RMNet.forward / memorize / segment compose them (batch 1)."""

    def __init__(self, gen, n, K, H, W, per_object_reader=False):
        self.gen, self.n, self.K, self.H, self.W = gen, n, K, H, W
        self.per_object_reader = per_object_reader   # run :155-160 one object at a time (p is 2 GB per object at 720p / T=40)
        self.h, self.w = (H + 15) // 16, (W + 15) // 16
        self.keys = self.vals = None

    def memorize(self, f):
        n, K, H, W, h, w = self.n, self.K, self.H, self.W, self.h, self.w
        dev = f["mask"].device
        masks = pad16(f["mask"][None], H, W)                                         # :212
        att, bb = self.gen.forward(masks.contiguous(), 0.5, 10, 64)                  # :244 -> the reference CUDA kernel
        k4 = torch.zeros(1, K, 128, 1, h, w, device=dev)
        v4 = torch.zeros(1, K, 512, 1, h, w, device=dev)
        k4[0, 1:n + 1, :, 0] = f["k4"]                                               # pad_memory :191-205
        v4[0, 1:n + 1, :, 0] = f["v4"]
        a16 = F.interpolate(att, scale_factor=1 / 16)[:, :, None, None]              # :245-246
        return k4 * a16, v4 * a16, bb                                                # :247-248

    def commit(self, f):
        k, v, _ = self.memorize(f)
        self.keys = k if self.keys is None else torch.cat([self.keys, k], dim=3)     # :420-421, :424-426
        self.vals = v if self.vals is None else torch.cat([self.vals, v], dim=3)

    def step(self, cur):
        """-> (mem_val [n,1024,h,w], prev_bbox [1,K,4], curr_bbox [1,K,4])"""
        n, H, W, h, w = self.n, self.H, self.W, self.h, self.w
        k, v, prev_bb = self.memorize(cur)
        this_keys = k if self.keys is None else torch.cat([self.keys, k], dim=3)     # :416-421
        this_vals = v if self.vals is None else torch.cat([self.vals, v], dim=3)
        warped, _ = torch_warp(cur["mask"][None], cur["flow"][None])                 # :284
        att, cur_bb = self.gen.forward(warped.contiguous(), 0.5, 10, 64)             # :286
        att = pad16(att, H, W)                                                       # :307
        a16 = F.interpolate(att[0, 1:n + 1, None], scale_factor=1 / 16)              # :330, :356
        k4e = cur["qk"][None].expand(n, -1, -1, -1) * a16                            # :332, :357
        v4e = cur["qv"][None].expand(n, -1, -1, -1) * a16                            # :333, :358
        m_key, m_val = this_keys[0, 1:n + 1].contiguous(), this_vals[0, 1:n + 1].contiguous()    # :348-349
        M, N = m_key.shape[2] * h * w, h * w
        if self.per_object_reader:
            mems = []
            for o in range(n):
                mi = torch.transpose(m_key[o:o + 1].view(1, 128, M), 1, 2)
                p = torch.softmax(torch.bmm(mi, k4e[o:o + 1].reshape(1, 128, N)) / math.sqrt(128), dim=1)
                mems.append(torch.bmm(m_val[o:o + 1].view(1, 512, M), p).view(1, 512, h, w))
                del p
            mem = torch.cat(mems, dim=0)
        else:
            mi = torch.transpose(m_key.view(n, 128, M), 1, 2)                        # :151-152
            p = torch.softmax(torch.bmm(mi, k4e.reshape(n, 128, N)) / math.sqrt(128), dim=1)   # :155-157
            mem = torch.bmm(m_val.view(n, 512, M), p).view(n, 512, h, w)             # :158-161
        return torch.cat([mem, v4e], dim=1), prev_bb, cur_bb                         # :163


def _pad_amounts(H, W, d=16):
    """utils/helpers.py:105-124 -> (lw, uw, lh, uh)"""
    nh, nw = (H + d - 1) // d * d, (W + d - 1) // d * d
    lh, lw = (nh - H) // 2, (nw - W) // 2
    return lw, nw - W - lw, lh, nh - H - lh


def torch_mask_epilogue(x, K, H, W, modes, new_mask):
    """models/rmnet.py:368-380, :289-302, :436-450 with torch's own ops on x's device (what the reference runs there).
    modes: 0 keep, 1 absent (:448), 2 new object (:442)."""
    n = x.shape[0]
    ps = F.softmax(x, dim=1)[:, 1]
    em = torch.zeros(1, K, *ps.shape[1:], device=x.device)
    em[0, 0] = torch.prod(1 - ps, dim=0)
    em[0, 1:n + 1] = ps
    em = torch.clamp(em, 1e-7, 1 - 1e-7)
    logit = torch.log((em / (1 - em)))
    lw, uw, lh, uh = _pad_amounts(H, W)
    logit = logit[:, :, lh:lh + H, lw:lw + W].clone()
    for j in range(K):
        if modes[j] == 2:
            logit[0, j] = new_mask[j].float() * 32.0605 - 16.1181
        if modes[j] == 1:
            logit[0, j] = -16.1181
    return logit, F.softmax(logit, dim=1)
