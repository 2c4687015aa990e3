"""CUDA-graph replay of the forward scan over sub-actions (no tracing compiler: the C ABI never allocates or
synchronises, so a whole `lax.scan(self.simulator.step_jax, state, actions)` (mpm_env.py:141) is capturable as is).

At the reference's own scene sizes (whip_rope: 67 particles, pour_water: 702) a step is launch-bound: S substeps x
~6 small kernels.  One graph launch replaces T x S x 6 launches; at the BASELINE sizes the kernels dominate and the
graph changes nothing.  GraphedMPMScan is forward only (evaluation rollouts, `env.step` without gradients);
GraphedMPMScanGrad captures the scan AND its adjoint (one graph each) behind a torch.autograd.Function, for the
differentiated `step_diff` of the launch-bound tasks (shape_rope: 30 sub-actions x 133 substeps per env step).
"""
import torch

from .mpm_simulator import _PRIM_LEAVES, _STATE_LEAVES, MPMState, PrimitiveState, _f32c, flatten_state


def _clone_state(state: MPMState) -> MPMState:
    prims = [PrimitiveState(*[t.detach().clone() for t in p]) for p in state.primitives]
    vals = {k: getattr(state, k).detach().clone() for k in state._fields if k != "primitives"}
    return state._replace(primitives=prims, **vals)


def _copy_state_(dst: MPMState, src: MPMState):
    for k in dst._fields:
        if k == "primitives":
            for pd, ps in zip(dst.primitives, src.primitives):
                for td, ts in zip(pd, ps):
                    td.copy_(ts)
        else:
            getattr(dst, k).copy_(getattr(src, k))


class GraphedMPMScan:
    """graph = GraphedMPMScan(sim, state, actions)   # actions [T, B, 6*n_prim]: captures T step_jax calls
    new_state = graph(state, actions)                  # copies into the static inputs, one cudaGraphLaunch

    The returned state aliases the graph's static output buffers: clone it if it must survive the next replay."""

    def __init__(self, sim, state: MPMState, actions: torch.Tensor):
        self.sim = sim
        self.s_in = _clone_state(state)
        self.a_in = actions.detach().clone()
        cur = torch.cuda.current_stream(sim.device)
        side = torch.cuda.Stream(sim.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():      # warm-up: workspaces and kernel attributes exist before capture
            self._scan(self.s_in, self.a_in)
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.s_out = self._scan(self.s_in, self.a_in)
        # the graph holds the raw pointer of the simulator's forward workspace: pin it, so that a later eager call that
        # needs a larger workspace fails loudly instead of freeing memory the next replay would write
        sim._ws_fwd.pins += 1
        self._pinned = True

    def close(self):
        """Drops the graph and releases its hold on the simulator's workspace."""
        if getattr(self, "_pinned", False):
            self.sim._ws_fwd.pins -= 1
            self._pinned = False
        self.graph = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _scan(self, state, actions):
        for t in range(actions.shape[0]):
            state, _ = self.sim.step_jax(state, actions[t])
        return state

    def __call__(self, state: MPMState, actions: torch.Tensor) -> MPMState:
        with torch.no_grad():
            _copy_state_(self.s_in, state)
            self.a_in.copy_(actions)
        if self.graph is None:
            raise RuntimeError("GraphedMPMScan was closed")
        self.graph.replay()
        return self.s_out


class _GraphScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scan, actions, *leaves):
        with torch.no_grad():
            for dst, src in zip(scan.leaves_in, leaves):
                dst.copy_(src)
            scan.a_in.copy_(actions)
        scan.fwd_graph.replay()
        ctx.scan = scan
        return tuple(t.clone() for t in scan.leaves_out)

    @staticmethod
    def backward(ctx, *gouts):
        scan = ctx.scan
        with torch.no_grad():
            for dst, g in zip(scan.gout, gouts):
                if g is None:
                    dst.zero_()
                else:
                    dst.copy_(g)
        scan.bwd_graph.replay()
        return (None, torch.stack([g for g in scan.gact]).clone(), *[g.clone() for g in scan.gin])


class GraphedMPMScanGrad:
    """Differentiable `lax.scan(simulator.step_jax, state, actions)` (mpm_env.py:141) as TWO CUDA graphs: the T forward
    calls (each step's input kept as its checkpoint in static buffers) and the T adjoint calls in reverse order.

        scan = GraphedMPMScanGrad(sim, state, actions)      # actions [T, B, 6*n_prim]
        new_state = scan(state, actions)                      # differentiable w.r.t. the state leaves and the actions

    The loss must be a function of the returned (final) state -- what the env's step_diff uses; the per-sub-action
    `ys` of the scan are not exposed.  The simulator must not be used for other shapes while the scan is alive (its
    workspaces are baked into the graphs)."""

    def __init__(self, sim, state: MPMState, actions: torch.Tensor):
        if sim.env_groups != 1:
            raise ValueError("GraphedMPMScanGrad needs a simulator with env_groups=1 (one stream to capture)")
        self.sim, self.n_prim = sim, int(sim.conf.n_primitive)
        self.template = state
        self.leaves_in = [_f32c(t).clone() for t in flatten_state(state, self.n_prim)]
        self.softness = [_f32c(state.primitives[q].softness).clone() for q in range(self.n_prim)]
        self.a_in = _f32c(actions).clone()
        T = self.a_in.shape[0]
        cur = torch.cuda.current_stream(sim.device)
        side = torch.cuda.Stream(sim.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():      # warm-up: workspaces and kernel attributes exist before capture
            out, _ = sim._call_fwd(self.leaves_in, self.softness, self.a_in[0])
            sim._call_bwd(self.leaves_in, self.softness, self.a_in[0], [torch.zeros_like(t) for t in out])
        cur.wait_stream(side)
        torch.cuda.synchronize(sim.device)
        self.fwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd_graph), torch.no_grad():
            self.ckpt, lv = [], self.leaves_in
            for t in range(T):
                self.ckpt.append(lv)
                lv, _ = sim._call_fwd(lv, self.softness, self.a_in[t])
            self.leaves_out = lv
        self.gout = [torch.zeros_like(t) for t in self.leaves_out]
        self.bwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.bwd_graph), torch.no_grad():
            g, self.gact = self.gout, [None] * T
            for t in reversed(range(T)):
                g, self.gact[t] = sim._call_bwd(self.ckpt[t], self.softness, self.a_in[t], list(g))
            self.gin = g
        sim._ws_fwd.pins += 1
        sim._ws_bwd.pins += 1
        self._pinned = True

    def close(self):
        if getattr(self, "_pinned", False):
            self.sim._ws_fwd.pins -= 1
            self.sim._ws_bwd.pins -= 1
            self._pinned = False
        self.fwd_graph = self.bwd_graph = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, state: MPMState, actions: torch.Tensor) -> MPMState:
        if self.fwd_graph is None:
            raise RuntimeError("GraphedMPMScanGrad was closed")
        out = _GraphScanFn.apply(self, actions, *flatten_state(state, self.n_prim))
        vals = dict(zip(_STATE_LEAVES, out[:len(_STATE_LEAVES)]))
        prims, n0, npl = [], len(_STATE_LEAVES), len(_PRIM_LEAVES)
        for q, ps in enumerate(state.primitives):
            if q < self.n_prim:
                ps = ps._replace(**dict(zip(_PRIM_LEAVES, out[n0 + q * npl: n0 + (q + 1) * npl])))
            prims.append(ps)
        return state._replace(primitives=prims, **vals)
