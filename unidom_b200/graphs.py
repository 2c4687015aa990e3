"""CUDA-graph replay of the forward scan over sub-actions (no tracing compiler: the C ABI never allocates or
synchronises, so a whole `lax.scan(self.simulator.step_jax, state, actions)` (mpm_env.py:141) is capturable as is).

At the reference's own scene sizes (whip_rope: 67 particles, pour_water: 702) a step is launch-bound: S substeps x
~6 small kernels.  One graph launch replaces T x S x 6 launches; at the BASELINE sizes the kernels dominate and the
graph changes nothing.  Forward only (evaluation rollouts, `env.step` without gradients); the differentiated path
keeps its eager launches.
"""
import torch

from .mpm_simulator import MPMState, PrimitiveState


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
