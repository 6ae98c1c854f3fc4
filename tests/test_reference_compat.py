"""Drop-in compatibility of the mirror modules: checkpoint keys and call signatures.

Checkpoint compatibility: benerf_b200.optimize.Model(args).build_network(args)
must own the reference Graph's 59 tensors under the same names and shapes, so that `{iter:06d}.tar` checkpoints written
by train.py:443-455 load into either implementation (graph.load_state_dict, test.py:101).  No GPU needed: parameter
holders are plain nn.Modules; the arithmetic lives in the CUDA library."""
import os
import sys
import types
from argparse import Namespace

import pytest
import torch


def _args(channels=3, n_importance=64):
    return Namespace(dataset="BeNeRF_Unreal", channels=channels, N_samples=64, N_importance=n_importance, multires=10, multires_views=4,
                     i_embed=0, use_viewdirs=True, use_barf_c2f=False, ndc=True, traj="spline", num_interpolated_pose=19,
                     rgb_crf_net_hidden=0, rgb_crf_net_width=128, event_crf_net_hidden=0, event_crf_net_width=128, chunk=4096,
                     lrate=5e-4, pose_lrate=1e-3, transform_lrate=1e-6, rgb_crf_lrate=5e-4, event_crf_lrate=5e-4,
                     barf_c2f=[0.1, 0.5], netdepth=8, netwidth=256)


def _expected(channels, fine):
    shapes = {}
    nets = ["nerf"] + (["nerf_fine"] if fine else [])
    for net in nets:
        lin = {"pts_linears.0": (256, 63), "views_linears.0": (128, 283), "feature_linear": (256, 256), "alpha_linear": (1, 256),
               "rgb_linear": (channels, 128)}
        for i in range(1, 8):
            lin[f"pts_linears.{i}"] = (256, 319 if i == 5 else 256)
        for k, (o, i) in lin.items():
            shapes[f"{net}.{k}.weight"], shapes[f"{net}.{k}.bias"] = (o, i), (o,)
    shapes.update({"evt_knot_pose_se3.params.weight": (4, 6), "rgb_knot_pose_se3.params.weight": (4, 6), "transform.params.weight": (1, 6)})
    for crf, mlp in (("rgb_crf", "mlp_gray"), ("event_crf", "mlp_luminance")):
        shapes.update({f"{crf}.{mlp}.0.weight": (128, 1), f"{crf}.{mlp}.0.bias": (128,), f"{crf}.{mlp}.2.weight": (1, 128), f"{crf}.{mlp}.2.bias": (1,)})
    return shapes


@pytest.mark.parametrize("channels,n_importance", [(3, 64), (1, 64), (1, 0)])
def test_state_dict_keys_and_shapes(channels, n_importance, monkeypatch):
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)       # build_network moves to cuda when it can
    from benerf_b200 import optimize
    args = _args(channels, n_importance)
    graph = optimize.Model(args).build_network(args)
    got = {k: tuple(v.shape) for k, v in graph.state_dict().items()}
    want = _expected(channels, n_importance > 0)
    assert got == want, set(got) ^ set(want)
    if channels == 3 and n_importance > 0:
        assert len(got) == 59
    # the five optimisers of model/optimize.py:36-55, in the reference's return order
    model = optimize.Model(args)
    model.build_network(args)
    opts = model.setup_optimizer(args)
    assert len(opts) == 5 and all(isinstance(o, torch.optim.Adam) for o in opts)
    assert [o.param_groups[0]["lr"] for o in opts] == [args.lrate, args.pose_lrate, args.transform_lrate, args.rgb_crf_lrate, args.event_crf_lrate]


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree not present (GPU box)")
def test_state_dict_round_trips_with_the_reference(monkeypatch):
    """Live check in the build container: the reference's own Graph loads our state dict (strict) and vice versa."""
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    for name in ("h5py", "hdf5plugin", "imageio", "imageio.v3"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.File = m.imwrite = m.imread = None
            monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.syspath_prepend("/root/reference")
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.") or k in ("spline", "run_nerf_helpers", "utils", "loss")]:
        monkeypatch.delitem(sys.modules, k, raising=False)
    try:
        ref_optimize = __import__("model.optimize", fromlist=["Model"])
    except Exception as e:            # the reference needs packages this image may lack
        pytest.skip(f"reference import failed: {type(e).__name__}: {e}")
    from benerf_b200 import optimize
    args = _args()
    ours = optimize.Model(args).build_network(args)
    ref = ref_optimize.Model(args).build_network(args)
    ref_sd = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ours.load_state_dict(ref_sd, strict=True)
    ref.load_state_dict({k: v.detach().cpu() for k, v in ours.state_dict().items()}, strict=True)


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree not present (GPU box)")
def test_call_signatures_match_the_reference(monkeypatch):
    """every callable train.py / test.py / run_nerf_helpers.py reach keeps its parameter names and order
    (extra keyword-only / defaulted parameters of the mirror are allowed at the end)."""
    import inspect
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    for name in ("h5py", "hdf5plugin", "imageio", "imageio.v3"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.File = m.imwrite = m.imread = None
            monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.syspath_prepend("/root/reference")
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.") or k in ("spline", "run_nerf_helpers", "utils", "loss")]:
        monkeypatch.delitem(sys.modules, k, raising=False)
    try:
        ref_nerf = __import__("model.nerf", fromlist=["Graph"])
        ref_opt = __import__("model.optimize", fromlist=["Model"])
        ref_spline = __import__("spline")
        ref_helpers = __import__("run_nerf_helpers")
    except Exception as e:
        pytest.skip(f"reference import failed: {type(e).__name__}: {e}")
    from benerf_b200 import nerf, optimize, spline, run_nerf_helpers

    def names(fn):
        return [p for p in inspect.signature(fn).parameters]

    pairs = [(ref_nerf.Graph.forward, nerf.Graph.forward), (ref_nerf.Graph.render, nerf.Graph.render),
             (ref_nerf.Graph.render_video, nerf.Graph.render_video),
             (ref_opt.Graph.get_pose_evt, optimize.Graph.get_pose_evt), (ref_opt.Graph.get_pose_rgb, optimize.Graph.get_pose_rgb),
             (ref_opt.Model.build_network, optimize.Model.build_network), (ref_opt.Model.setup_optimizer, optimize.Model.setup_optimizer),
             (ref_spline.cubic_spline_pose_unit_time, spline.cubic_spline_pose_unit_time),
             (ref_spline.linear_pose_unit_time, spline.linear_pose_unit_time),
             (ref_helpers.init_nerf, run_nerf_helpers.init_nerf),
             (ref_helpers.render_video_test, run_nerf_helpers.render_video_test),
             (ref_helpers.render_image_test, run_nerf_helpers.render_image_test)]
    for ref_fn, our_fn in pairs:
        r, o = names(ref_fn), names(our_fn)
        assert o[:len(r)] == r, (ref_fn.__qualname__, r, o)
        for extra in o[len(r):]:
            assert inspect.signature(our_fn).parameters[extra].default is not inspect.Parameter.empty, (our_fn.__qualname__, extra)
