"""Drop-in replacement of the resolution regressor (reference model/res_regressor.py:106-177).

`MPCResRgrNoPool(config)` keeps the reference's constructor, its `state_dict` layout (`model.0.weight` ...
`model.19.bias`, so `load_state_dict(torch.load(...))` of a reference checkpoint works, env/flex_env.py:986-989),
`forward(x)` and `infer_param(init_img, goal_img) -> int`.  The network runs in libpilegnn (csrc/rgr.cu: five
stride-2 convolutions as split-K implicit GEMMs, five linear layers as weight-streaming GEMVs), the batch-1 call
is captured as one CUDA graph.  The input planes are built on the host with the same OpenCV calls as the reference
(distance transform + INTER_AREA resize, :153-172) -- third-party arithmetic that is kept, not re-implemented,
like the goal shaping of the reward.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops

CONV_CHANNELS = (6, 64, 128, 256, 512, 512)
FC_WIDTHS = (512 * 7 * 7, 4096, 1024, 256, 64, 1)


def regressor_input(init_img, goal_img, state_h, state_w):
    """Two binary [H,W] masks -> float32 [6, state_h, state_w] (res_regressor.py:153-172)."""
    import cv2
    assert init_img.shape == goal_img.shape
    init_img_dist = cv2.distanceTransform((1 - init_img).astype(np.uint8), cv2.DIST_L2, 5) / (init_img.shape[0])
    goal_img_dist = cv2.distanceTransform((1 - goal_img).astype(np.uint8), cv2.DIST_L2, 5) / (goal_img.shape[0])
    init_exclude_goal = np.logical_and(init_img, 1 - goal_img).astype(np.float32)
    goal_exclude_init = np.logical_and(goal_img, 1 - init_img).astype(np.float32)
    planes = [init_img, goal_img, init_img_dist, goal_img_dist, init_exclude_goal, goal_exclude_init]
    return np.stack([cv2.resize(p, (state_w, state_h), interpolation=cv2.INTER_AREA) for p in planes], axis=0).astype(np.float32)


class MPCResRgrNoPool(nn.Module):

    def __init__(self, config):
        super(MPCResRgrNoPool, self).__init__()
        self.config = config
        self.state_h = config['train_res_cls']['state_h']
        self.state_w = config['train_res_cls']['state_w']
        self.res_dim = config['train_res_cls']['res_dim']
        # same construction order as the reference, so torch.manual_seed(s) reproduces its initial weights and the
        # Sequential indices (0, 2, 4, 6, 8 | 11, 13, 15, 17, 19) give the checkpoint's keys
        layers = []
        for cin, cout in zip(CONV_CHANNELS[:-1], CONV_CHANNELS[1:]):
            layers += [nn.Conv2d(cin, cout, 4, 2, 1), nn.LeakyReLU(negative_slope=0.2)]
        layers.append(nn.Flatten())
        for k, (fin, fout) in enumerate(zip(FC_WIDTHS[:-1], FC_WIDTHS[1:])):
            layers.append(nn.Linear(fin, fout))
            if k < len(FC_WIDTHS) - 2:
                layers.append(nn.LeakyReLU(negative_slope=0.2))
        self.model = nn.Sequential(*layers)
        self._packed = None
        self._graph = None

    # ---- weights -> one float buffer in state_dict order, re-packed only when a parameter changed ----------------
    def packed_params(self, device):
        params = [p for _, p in self.named_parameters()]
        stamp = tuple((p.data_ptr(), p._version) for p in params) + (str(device),)
        if self._packed is None or self._packed[0] != stamp:
            lib = _lib.load()
            total = lib.pile_rgr_param_offset(2 * 10)
            buf = torch.empty(total, dtype=torch.float32, device=device)
            for i, p in enumerate(params):
                off = lib.pile_rgr_param_offset(i)
                if off + p.numel() != lib.pile_rgr_param_offset(i + 1):
                    raise _lib.PileLibraryError("regressor tensor %d has %d elements, library expects %d" %
                                                (i, p.numel(), lib.pile_rgr_param_offset(i + 1) - off))
                buf[off:off + p.numel()].copy_(p.detach().reshape(-1))
            self._packed = (stamp, buf)
            self._graph = None
        return self._packed[1]

    def forward(self, x):
        """x [B, 6, state_h, state_w] CUDA float -> [B, 1] (inference only: no autograd through this path)."""
        x = ops._f32(x.detach())
        ops._require_cuda(x, "x")
        B, C, Hh, Ww = x.shape
        assert C == CONV_CHANNELS[0]
        lib = _lib.load()
        params = self.packed_params(x.device)
        nbytes = lib.pile_rgr_workspace_bytes(B, Hh, Ww)
        if nbytes < 0:
            raise _lib.PileLibraryError("regressor input %dx%d does not reduce to 7x7" % (Hh, Ww))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        y = torch.empty(B, 1, dtype=torch.float32, device=x.device)
        _lib.check(lib.pile_rgr_forward(_lib.ptr(params), _lib.ptr(x), B, Hh, Ww, _lib.ptr(ws), _lib.ptr(y), ops._stream()),
                   "pile_rgr_forward")
        return y

    def _graph_forward(self, x_host):
        """Batch-1 call with static buffers, captured once: H2D of the 6 planes -> 15 launches -> D2H of one float."""
        dev = torch.device('cuda')
        lib = _lib.load()
        params = self.packed_params(dev)
        g = self._graph
        if g is None or g['shape'] != tuple(x_host.shape):
            _, Hh, Ww = x_host.shape
            nbytes = lib.pile_rgr_workspace_bytes(1, Hh, Ww)
            if nbytes < 0:
                raise _lib.PileLibraryError("regressor input %dx%d does not reduce to 7x7" % (Hh, Ww))
            g = {'shape': tuple(x_host.shape), 'x': torch.zeros((1,) + tuple(x_host.shape), device=dev),
                 'ws': torch.empty(nbytes, dtype=torch.uint8, device=dev), 'y': torch.zeros(1, 1, device=dev),
                 'pin': torch.empty((1,) + tuple(x_host.shape), dtype=torch.float32).pin_memory(),
                 'out': torch.empty(1, 1, dtype=torch.float32).pin_memory()}

            def enqueue():
                _lib.check(lib.pile_rgr_forward(_lib.ptr(params), _lib.ptr(g['x']), 1, Hh, Ww, _lib.ptr(g['ws']),
                                                _lib.ptr(g['y']), ops._stream()), "pile_rgr_forward")
            enqueue()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                enqueue()
            g['graph'] = graph
            self._graph = g
        g['pin'].copy_(torch.from_numpy(x_host[None]))
        g['x'].copy_(g['pin'], non_blocking=True)
        g['graph'].replay()
        g['out'].copy_(g['y'], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(g['out'][0, 0])

    def infer_param(self, init_img, goal_img):
        # init / goal: binary numpy of shape (H, W)  (res_regressor.py:146-177)
        assert init_img.shape == goal_img.shape
        x = regressor_input(init_img, goal_img, self.state_h, self.state_w)
        particle_num = self._graph_forward(x)
        return int(particle_num)
