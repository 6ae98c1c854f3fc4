"""Mirror of model/optimize.py: the concrete Model / Graph (knots, transform, optimisers, poses)."""
import torch

from . import nerf
from .component import ColorToneMapper, LuminanceToneMapper, ControlKnotLieAlgebra, TransformationLieAlgebra
from .spline import SplineFn


class Model(nerf.Model):
    def __init__(self, args):
        self.graph = Graph(args, D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)

    def build_network(self, args, poses=None, event_poses=None):
        """model/optimize.py:11-34: knots = rand(4,6) * 0.01, transform = 0, CRF holders."""
        g = self.graph
        g.evt_knot_pose_se3 = ControlKnotLieAlgebra(4)
        g.rgb_knot_pose_se3 = ControlKnotLieAlgebra(4)          # allocated, never used upstream (Q6); kept for the state dict
        g.transform = TransformationLieAlgebra(1)
        g.rgb_crf = ColorToneMapper(hidden=args.rgb_crf_net_hidden, width=args.rgb_crf_net_width, input_type="Gray")
        g.event_crf = LuminanceToneMapper(hidden=args.event_crf_net_hidden, width=args.event_crf_net_width, input_type="Gray")
        parm_evt = torch.cat([torch.rand(1, 6) * 0.01 for _ in range(4)])
        g.evt_knot_pose_se3.params.weight.data = torch.nn.Parameter(parm_evt)
        g.transform.params.weight.data = torch.nn.Parameter(torch.zeros(1, 6))
        g.rgb_crf.weights_biases_init()
        g.event_crf.weights_biases_init()
        if torch.cuda.is_available():
            g.to("cuda")        # the reference runs with default tensor type cuda (train.py:472)
        return g

    def setup_optimizer(self, args):
        """model/optimize.py:36-55: five Adam optimisers, same grouping and return order."""
        g = self.graph
        grad_vars = list(g.nerf.parameters())
        if args.N_importance > 0:
            grad_vars += list(g.nerf_fine.parameters())
        self.optim_nerf = torch.optim.Adam(params=grad_vars, lr=args.lrate)
        self.optim_pose = torch.optim.Adam(params=list(g.evt_knot_pose_se3.parameters()), lr=args.pose_lrate)
        self.optim_transform = torch.optim.Adam(params=list(g.transform.parameters()), lr=args.transform_lrate)
        self.optim_event_crf = torch.optim.Adam(params=list(g.event_crf.mlp_luminance.parameters()), lr=args.event_crf_lrate)
        self.optim_rgb_crf = torch.optim.Adam(params=list(g.rgb_crf.mlp_gray.parameters()), lr=args.rgb_crf_lrate)
        return self.optim_nerf, self.optim_pose, self.optim_transform, self.optim_rgb_crf, self.optim_event_crf


class Graph(nerf.Graph):
    def _poses(self, args, ts2, num, with_transform):
        eng = self.engine(args)
        # built on the device: a host linspace + .to(device) is a pageable copy, i.e. a stream synchronisation every iteration
        ts = torch.linspace(float(ts2[0]), float(ts2[1]), num, device=eng.device)
        if args.traj not in ("linear", "spline"):
            raise ValueError(args.traj)
        return SplineFn.apply(eng, args.traj, ts, self.evt_knot_pose_se3.params.weight,
                               self.transform.params.weight if with_transform else None)

    def get_pose_evt(self, args, events_ts, seg_num=None):
        """model/optimize.py:58-82: P = 2 (window start/end) unless seg_num is given."""
        return self._poses(args, events_ts, 2 if seg_num is None else seg_num, with_transform=False)

    def get_pose_rgb(self, args, exposure_ts, seg_num=None):
        """model/optimize.py:84-111: RGB knots = event knots + transform, added in se(3) (Q6)."""
        return self._poses(args, exposure_ts, args.num_interpolated_pose if seg_num is None else seg_num, with_transform=True)
