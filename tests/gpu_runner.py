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
