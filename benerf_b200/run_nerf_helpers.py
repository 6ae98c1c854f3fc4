"""Mirror of the run_nerf_helpers.py entry points the callers use around the hot path."""
import torch


def init_weights(linear):
    """run_nerf_helpers.py:194-197: Xavier-uniform weight, zero bias."""
    torch.nn.init.xavier_uniform_(linear.weight)
    torch.nn.init.zeros_(linear.bias)


def init_nerf(nerf):
    """run_nerf_helpers.py:199-208."""
    for linear_pt in nerf.pts_linears:
        init_weights(linear_pt)
    for linear_view in nerf.views_linears:
        init_weights(linear_view)
    init_weights(nerf.feature_linear)
    init_weights(nerf.alpha_linear)
    init_weights(nerf.rgb_linear)


# ---- evaluation drivers (run_nerf_helpers.py:117-171): same signatures, Graph.render_video underneath -------------------
def to8bit(x):
    """utils/img_utils.py:19-20."""
    import numpy as np
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def _write_png(path, img8):
    """imageio.v3.imwrite stand-in (the reference's only use of imageio on this path): 8-bit gray or RGB PNG."""
    import struct
    import zlib
    import numpy as np
    img8 = np.ascontiguousarray(img8)
    if img8.ndim == 2:
        img8 = img8[:, :, None]
    h, w, c = img8.shape
    raw = b"".join(b"\x00" + img8[r].tobytes() for r in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if c == 1 else 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


@torch.no_grad()
def render_video_test(iter_step, graph, render_poses, H, W, K, args, remap):
    """run_nerf_helpers.py:117-140: one full-image render per pose -> (rgbs [P,H,W,C], disps [P,H,W]) numpy."""
    import numpy as np
    rgbs, disps = [], []
    for pose in render_poses:
        ret = graph.render_video(iter_step, pose[None, :3, :4], H, W, K, args, remap, type="rgb")
        if getattr(args, "optimize_rgb_crf", False):                     # run_nerf_helpers.py:125-126
            ret["rgb_map"] = graph.rgb_crf.forward(ret["rgb_map"])
        rgbs.append(ret["rgb_map"].cpu().numpy())
        disps.append(ret["disp_map"].cpu().numpy())
    return np.stack(rgbs, 0), np.stack(disps, 0)


@torch.no_grad()
def render_image_test(iter_step, graph, render_poses, H, W, K, args, logdir, remap, dir=None, need_depth=True):
    """run_nerf_helpers.py:142-171: renders every pose, writes 8-bit PNGs under logdir/dir/img_test_{iter:06d}/."""
    import os
    import numpy as np
    img_dir = os.path.join(logdir, dir, "img_test_{:06d}".format(iter_step))
    os.makedirs(img_dir, exist_ok=True)
    imgs, depth = [], []
    for j, pose in enumerate(render_poses):
        ret = graph.render_video(iter_step, pose[None, :3, :4], H, W, K, args, remap, type="rgb")
        if getattr(args, "optimize_rgb_crf", False):                     # run_nerf_helpers.py:152-153
            ret["rgb_map"] = graph.rgb_crf.forward(ret["rgb_map"])
        rgb8 = to8bit(ret["rgb_map"].cpu().numpy())
        _write_png(os.path.join(img_dir, dir[11:] + "{:03d}.png".format(j)), rgb8.squeeze())
        imgs.append(rgb8)
        if need_depth:
            depths = ret["disp_map"].cpu().numpy()
            depth8 = to8bit(depths / np.max(depths))
            _write_png(os.path.join(img_dir, "depth_{:03d}.png".format(j)), depth8)
            depth.append(depth8)
    return imgs, depth
