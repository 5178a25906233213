"""Helpers that run the product's picture-level entries on the GPU through the C ABI (ctypes) with torch used
only to own device memory."""
import ctypes as C

import numpy as np
import torch

import common as cm
import svtb200 as sb


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_gpu_me(params, src, refs, stream=None):
    """src / refs: tuples of 3 numpy planes. Returns a MeBuffers filled from the GPU result."""
    lib = sb.load()
    n_sb = ((params.full.width + 63) // 64) * ((params.full.height + 63) // 64)
    d_src = [dev(x) for x in src]
    d_refs = [[dev(x) for x in r] for r in refs]
    s = sb.MePlanes(*[t.data_ptr() for t in d_src])
    r = (sb.MePlanes * 8)(*[sb.MePlanes(*[t.data_ptr() for t in rr]) for rr in d_refs])
    out = cm.MeBuffers(n_sb)
    d_out = {k: torch.zeros(v.shape if v.dtype.fields is None else (v.size * v.dtype.itemsize,),
                            dtype=torch.uint8 if v.dtype.fields is not None else getattr(torch, str(v.dtype)),
                            device="cuda")
             for k, v in out.fields().items()}
    o = sb.MeOutputs(d_out["best_sad"].data_ptr(), d_out["best_mv"].data_ptr(), d_out["hme"].data_ptr(),
                     d_out["me_mv"].data_ptr(), d_out["me_cand"].data_ptr(), d_out["total_cand"].data_ptr(),
                     d_out["rc"].data_ptr())
    scratch = torch.zeros(lib.svt_b200_me_scratch_bytes(C.byref(params)), dtype=torch.uint8, device="cuda")
    sb.check(lib.svt_b200_me_picture(C.byref(params), C.byref(s), r, C.byref(o), scratch.data_ptr(), None), lib)
    torch.cuda.synchronize()
    for k, v in out.fields().items():
        host = d_out[k].cpu().numpy()
        if v.dtype.fields is not None:
            v[...] = host.view(v.dtype).reshape(v.shape)
        else:
            v[...] = host
    return out


class DevYuv:
    """Device copy of a common.Yuv (same padded layout)."""

    def __init__(self, yuv):
        self.yuv = yuv
        self.t = [torch.from_numpy(b.view(np.int16) if b.dtype == np.uint16 else b).cuda() for b in yuv.bufs]

    def struct(self):
        y = self.yuv
        p = y.pad
        ptrs = [t.data_ptr() + (p * b.shape[1] + p) * b.itemsize for t, b in zip(self.t, y.bufs)]
        return sb.Frame(ptrs[0], ptrs[1], ptrs[2], y.bufs[0].shape[1], y.bufs[1].shape[1], y.w, y.h, y.bd)

    def download(self):
        o = self.yuv.copy()
        for b, t in zip(o.bufs, self.t):
            h = t.cpu().numpy()
            b[...] = h.view(np.uint16) if b.dtype == np.uint16 else h
        return o


def run_gpu_cdef_search(p, rec, src, skip):
    lib = sb.load()
    nfb = ((p.mi_rows + 15) // 16) * ((p.mi_cols + 15) // 16)
    dr, ds = DevYuv(rec), DevYuv(src)
    dskip = dev(skip)
    mse = torch.zeros(2 * nfb * 64, dtype=torch.int64, device="cuda")
    rs, ss = dr.struct(), ds.struct()
    sb.check(lib.svt_b200_cdef_search(C.byref(p), C.byref(rs), C.byref(ss), C.c_void_p(dskip.data_ptr()),
                                      skip.shape[1], C.c_void_p(mse.data_ptr()), None), lib)
    torch.cuda.synchronize()
    return mse.cpu().numpy().view(np.uint64).reshape(2, nfb, 64)


def run_gpu_cdef_apply(p, rec, skip, idx):
    lib = sb.load()
    dr, do = DevYuv(rec), DevYuv(rec.copy())
    dskip, didx = dev(skip), dev(idx)
    rs, os_ = dr.struct(), do.struct()
    sb.check(lib.svt_b200_cdef_apply(C.byref(p), C.byref(rs), C.byref(os_), C.c_void_p(dskip.data_ptr()), skip.shape[1],
                                     C.c_void_p(didx.data_ptr()), None), lib)
    torch.cuda.synchronize()
    return do.download()


def run_gpu_encode_tus(p, src, pred, tus, with_cul=False):
    lib = sb.load()
    ts = p.tx_size
    w, h = sb.TX_W[ts], sb.TX_H[ts]
    n = min(w, 32) * min(h, 32)
    ds, dp, dr = DevYuv(src), DevYuv(pred), DevYuv(pred.copy())
    arr = (sb.Tu * len(tus))(*tus)
    dt = torch.from_numpy(np.frombuffer(arr, dtype=np.int32).copy()).cuda()
    q = torch.zeros(len(tus) * n, dtype=torch.int32, device="cuda")
    eob = torch.zeros(len(tus), dtype=torch.int16, device="cuda")
    cul = torch.full((len(tus),), -1, dtype=torch.int32, device="cuda")
    ss, ps, rs = ds.struct(), dp.struct(), dr.struct()
    sb.check(lib.svt_b200_encode_tus_cul(C.byref(p), C.byref(ss), C.byref(ps), C.byref(rs), C.c_void_p(dt.data_ptr()), len(tus),
                                         C.c_void_p(q.data_ptr()), C.c_void_p(eob.data_ptr()), C.c_void_p(cul.data_ptr()), None), lib)
    torch.cuda.synchronize()
    out = (dr.download(), q.cpu().numpy().reshape(len(tus), n), eob.cpu().numpy().view(np.uint16))
    return out + (cul.cpu().numpy(),) if with_cul else out


def run_gpu_encode_tus_ex(qsets, use_fp, src, pred, tus):
    """svt_b200_encode_tus_ex on a HOST list of sb.TuEx; returns (recon, [levels per unit], eob, cul_level)."""
    lib = sb.load()
    ds, dp, dr = DevYuv(src), DevYuv(pred), DevYuv(pred.copy())
    arr = (sb.TuEx * len(tus))(*tus)
    qs = ((sb.QuantPlane * 3) * len(qsets))()
    for i, three in enumerate(qsets):
        for pl in range(3):
            qs[i][pl] = three[pl]
    p = sb.EncodeParamsEx(use_fp, len(qsets), qs)
    q = torch.full((max(1, len(tus)) * 1024,), 77, dtype=torch.int32, device="cuda")
    eob = torch.zeros(max(1, len(tus)), dtype=torch.int16, device="cuda")
    cul = torch.full((max(1, len(tus)),), -1, dtype=torch.int32, device="cuda")
    scratch = torch.zeros(len(tus) * 28 + 256, dtype=torch.uint8, device="cuda")
    offs = (C.c_int64 * (len(tus) + 1))()
    ss, ps, rs = ds.struct(), dp.struct(), dr.struct()
    sb.check(lib.svt_b200_encode_tus_ex(C.byref(p), C.byref(ss), C.byref(ps), C.byref(rs), arr, len(tus), C.c_void_p(q.data_ptr()), offs,
                                        C.c_void_p(eob.data_ptr()), C.c_void_p(cul.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                        scratch.numel(), None), lib)
    torch.cuda.synchronize()
    qh = q.cpu().numpy()
    levels = [qh[offs[i]:offs[i + 1]] for i in range(len(tus))]
    return dr.download(), levels, eob.cpu().numpy().view(np.uint16)[:len(tus)], cul.cpu().numpy()[:len(tus)], qh[offs[len(tus)]:]


def run_gpu_dlf(p, frame, flat):
    lib = sb.load()
    df = DevYuv(frame.copy())
    dmi = torch.from_numpy(np.frombuffer(flat, dtype=np.uint8).copy()).cuda()
    st = df.struct()
    sb.check(lib.svt_b200_dlf_frame(C.byref(p), C.byref(st), C.c_void_p(dmi.data_ptr()), None), lib)
    torch.cuda.synchronize()
    return df.download()


def run_gpu_pick(p, rec, src, flat):
    lib = sb.load()
    dr, ds, dt = DevYuv(rec.copy()), DevYuv(src), DevYuv(rec.copy())
    dmi = torch.from_numpy(np.frombuffer(flat, dtype=np.uint8).copy()).cuda()
    scratch = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    rs, ss, ts = dr.struct(), ds.struct(), dt.struct()
    out = (C.c_int32 * 4)()
    sb.check(lib.svt_b200_pick_filter_level(C.byref(p), C.byref(rs), C.byref(ss), C.byref(ts), C.c_void_p(dmi.data_ptr()),
                                            C.c_void_p(scratch.data_ptr()), out, None), lib)
    torch.cuda.synchronize()
    return list(out), dr.download()


def run_gpu_lr(cdef, dblk, units, unit_sizes, frame_types, optimized):
    lib = sb.load()
    dc, dd = DevYuv(cdef), DevYuv(dblk)
    blank = cdef.copy()
    for b in blank.bufs:
        b[...] = 0
    do = DevYuv(blank)
    dus = [torch.from_numpy(np.frombuffer(u, dtype=np.uint8).copy()).cuda() for u in units]
    p = sb.LrFrameParams()
    for i in range(3):
        p.plane[i].frame_restoration_type = frame_types[i]
        p.plane[i].restoration_unit_size = unit_sizes[i]
        p.plane[i].units = dus[i].data_ptr()
    p.optimized_lr = optimized
    cs, ds, os_ = dc.struct(), dd.struct(), do.struct()
    sb.check(lib.svt_b200_lr_frame(C.byref(p), C.byref(cs), C.byref(ds), C.byref(os_), None), lib)
    torch.cuda.synchronize()
    return do.download()


def run_gpu_sse(a, b):
    lib = sb.load()
    da, db = DevYuv(a), DevYuv(b)
    out = torch.zeros(3, dtype=torch.int64, device="cuda")
    sa, sbb = da.struct(), db.struct()
    sb.check(lib.svt_b200_frame_sse(C.byref(sa), C.byref(sbb), C.c_void_p(out.data_ptr()), None), lib)
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint64)


def run_gpu_inter_predict(refs, pred, jobs, scratch_bytes=None):
    """refs: list of common.Yuv, pred: common.Yuv (written), jobs: numpy record array (sb.INTER_JOB_DTYPE)."""
    lib = sb.load()
    lib.svt_b200_inter_predict_scratch_bytes.restype = C.c_size_t
    if scratch_bytes is None:
        scratch_bytes = lib.svt_b200_inter_predict_scratch_bytes(len(jobs), pred.w, pred.h)
    scratch = torch.zeros(scratch_bytes, dtype=torch.uint8, device="cuda")
    d_refs = [DevYuv(r) for r in refs]
    d_pred = DevYuv(pred)
    arr = (sb.Frame * len(refs))(*[r.struct() for r in d_refs])
    ps = d_pred.struct()
    d_jobs = torch.from_numpy(np.ascontiguousarray(jobs).view(np.uint8)).cuda()
    sb.check(lib.svt_b200_inter_predict(arr, len(refs), C.byref(ps), C.c_void_p(d_jobs.data_ptr()), len(jobs),
                                        C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch_bytes), None), lib)
    torch.cuda.synchronize()
    return d_pred.download()
