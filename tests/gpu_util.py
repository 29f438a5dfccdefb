"""GPU-side helpers for the -m gpu tests: torch is used only for device memory and streams."""
import ctypes as C

import numpy as np

from dfb_b200 import capi


def to_dev(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def stream_handle(torch):
    return torch.cuda.current_stream().cuda_stream


def fuse_pool(torch, fuser, depth_pool, color_pool, out_hw, n=None):
    """depth_pool [F][N][H][W] u16, color_pool [F][N][H][W][3] u8 (numpy) -> fused [F][Ho][Wo][3] (numpy)."""
    n = depth_pool.shape[0] if n is None else n
    d = to_dev(torch, depth_pool.view(np.int16))
    c = to_dev(torch, color_pool)
    out = torch.empty((n, out_hw[0], out_hw[1], 3), dtype=torch.uint8, device="cuda")
    fuser.fuse_device(d.data_ptr(), c.data_ptr(), out.data_ptr(), n, stream_handle(torch))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def synth_pool_gpu(torch, dfb, cfg, first, count):
    """Generate framesets first..first+count-1 on the device; returns (depth int16-view tensor, color tensor)."""
    depth = torch.empty((count, cfg.n_cameras, cfg.height, cfg.width), dtype=torch.int16, device="cuda")
    color = torch.empty((count, cfg.n_cameras, cfg.height, cfg.width, 3), dtype=torch.uint8, device="cuda")
    scenes = cfg.scenes(first, count)
    cams = cfg.cams_array()
    capi.check(dfb.lib().dfb_synth_generate(scenes, count, cams, cfg.n_cameras, depth.data_ptr(), color.data_ptr(),
                                            stream_handle(torch)))
    torch.cuda.synchronize()
    return depth, color


def u16(t):
    """int16 device tensor -> numpy uint16"""
    return t.cpu().numpy().view(np.uint16)
