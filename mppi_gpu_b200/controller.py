"""Host-side mirror of the reference controller object, over the C ABI.

`PointMassModel` keeps the reference's constructor and method names
(include/point_mass.hpp:23-44 of NicolayP/mppi_gpu): get_act, memcpy_set_data, set_x,
get_u, get_inf -- so tests and drivers read like the reference's main.cu loop
(src/main.cu:311-371).  All compute happens in libmppi_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class PointMassModel:
    """PointMassModel(nb_sim, steps, dt, state_dim, act_dim, verbose=False) as in the reference,
    plus keyword-only extensions for what the reference hard-codes or lacks."""

    def __init__(self, nb_sim, steps, dt, state_dim, act_dim, verbose=False, *, lam=1.0,
                 sigma=0.025, inv_sigma=1.0, init_act=0.0, max_act=1.0, seed=0, flags=0,
                 device=0, rank=0, world_size=1, comm_id=None, comm=None, devices=None,
                 state_gain=None, act_gain=None, philox_rounds=10):
        self._lib = capi.load()
        p = capi.MppiParams()
        capi.check(self._lib.mppi_params_default(C.byref(p)))
        p.samples, p.horizon, p.dt = int(nb_sim), int(steps), float(dt)
        p.state_dim, p.act_dim, p.verbose = int(state_dim), int(act_dim), int(bool(verbose))
        p.lambda_ = float(lam)
        for name, val in (("sigma", sigma), ("inv_sigma", inv_sigma), ("init_act", init_act),
                          ("max_act", max_act)):
            arr = np.broadcast_to(np.asarray(val, dtype=np.float32), (int(act_dim),)) \
                if int(act_dim) <= capi.MAX_ACT else np.zeros(0, np.float32)
            for i, v in enumerate(arr):
                getattr(p, name)[i] = float(v)
        p.seed, p.flags, p.device = int(seed), int(flags), int(device)
        p.philox_rounds = int(philox_rounds)
        if state_gain is not None or act_gain is not None:
            # MPPI_MODEL_LINEAR_AXIS: caller-given gains {g0,g1,g2,g3}, {b0,b1}
            assert len(state_gain) == 4 and len(act_gain) == 2
            p.model = capi.MODEL_LINEAR_AXIS
            for i in range(4):
                p.state_gain[i] = float(state_gain[i])
            for i in range(2):
                p.act_gain[i] = float(act_gain[i])
        p.rank, p.world_size = int(rank), int(world_size)
        if world_size > 1:
            p.comm = capi.COMM_NCCL if comm is None else int(comm)
            if p.comm == capi.COMM_NCCL:
                assert comm_id is not None and len(comm_id) == capi.COMM_ID_BYTES
                for i, b in enumerate(bytes(comm_id)):
                    p.comm_id[i] = b
        self.params = p
        self._h = C.c_void_p()
        if devices is not None:
            # one process driving several GPUs: K sharded over `devices`, peer-memory exchange
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            capi.check(self._lib.mppi_create_multi(C.byref(p), devs, len(devices), C.byref(self._h)))
        else:
            capi.check(self._lib.mppi_create(C.byref(p), C.byref(self._h)))
        self.K, self.T, self.S, self.A = int(nb_sim), int(steps), int(state_dim), int(act_dim)
        kl, ko = C.c_int64(), C.c_int64()
        capi.check(self._lib.mppi_local_samples(self._h, C.byref(kl), C.byref(ko)))
        self.k_local, self.k_offset = int(kl.value), int(ko.value)

    # ------------------------------------------------------------------ reference API
    def memcpy_set_data(self, x, u, goal, w):
        x, u, goal, w = capi.f32(x), capi.f32(u), capi.f32(goal), capi.f32(w)
        assert x.size == self.S and u.size == self.T * self.A
        assert goal.size == self.S and w.size == self.S
        capi.check(self._lib.mppi_set_problem(self._h, x.ctypes.data, u.ctypes.data,
                                              goal.ctypes.data, w.ctypes.data))

    def set_terminal_weights(self, w_final):
        """The final state gets a Cost object of its own (Cost::final_cost with w_final); None:
        back to the reference's single object.  No reference counterpart."""
        if w_final is None:
            capi.check(self._lib.mppi_set_terminal_weights(self._h, None))
            return
        wf = capi.f32(w_final)
        assert wf.size == self.S
        capi.check(self._lib.mppi_set_terminal_weights(self._h, wf.ctypes.data))

    def set_x(self, x):
        x = capi.f32(x)
        assert x.size == self.S
        capi.check(self._lib.mppi_set_state(self._h, x.ctypes.data))

    def set_q(self, q, q_dot):
        """The state as positions and velocities (README: "decouple x into q and q_dot"); the
        state vector is [q, q_dot] (src/point_mass_gpu.cu:97-106: x[i], x[i + S/2])."""
        q, q_dot = capi.f32(q), capi.f32(q_dot)
        assert q.size == self.A and q_dot.size == self.A
        self.set_x(np.concatenate([q, q_dot]))

    def get_act(self, next_act=None):
        out = np.zeros(self.A, np.float32) if next_act is None else next_act
        capi.check(self._lib.mppi_step(self._h, out.ctypes.data))
        return out

    def get_u(self):
        u = np.zeros((self.T, self.A), np.float32)
        capi.check(self._lib.mppi_get_u(self._h, u.ctypes.data))
        return u

    def get_inf(self, want_x=False, want_e=True):
        """Returns dict(x, u, e, cost, beta, nabla, weight) like the reference's get_inf
        (src/point_mass.cu:236-262); x only on request (debug recomputation)."""
        K, T, S, A = self.k_local, self.T, self.S, self.A
        x = np.zeros((K, T + 1, S), np.float32) if want_x else None
        e = np.zeros((K, T, A), np.float32) if want_e else None
        u = np.zeros((T, A), np.float32)
        cost = np.zeros(K, np.float32)
        weight = np.zeros(K, np.float32)
        beta, nabla = C.c_float(), C.c_float()
        capi.check(self._lib.mppi_get_info(
            self._h, x.ctypes.data if want_x else None, u.ctypes.data,
            e.ctypes.data if want_e else None, cost.ctypes.data, C.addressof(beta),
            C.addressof(nabla), weight.ctypes.data))
        return dict(x=x, u=u, e=e, cost=cost, beta=np.float32(beta.value),
                    nabla=np.float32(nabla.value), weight=weight)

    # ------------------------------------------------------------------ extensions
    def next(self, x):
        """ControllerBase::next(x) of the reference's intended interface
        (include/controller_base.hpp:9-17) == set_x(x); get_act()."""
        self.set_x(x)
        return self.get_act()

    def set_u(self, u):
        u = capi.f32(u)
        assert u.size == self.T * self.A
        capi.check(self._lib.mppi_set_u(self._h, u.ctypes.data))

    def set_noise(self, e):
        e = capi.f32(e)
        assert e.size == self.k_local * self.T * self.A
        capi.check(self._lib.mppi_set_noise(self._h, e.ctypes.data))

    def set_noise_mode(self, injected):
        capi.check(self._lib.mppi_set_noise_mode(self._h, int(bool(injected))))

    def sample_only(self, step):
        capi.check(self._lib.mppi_sample_only(self._h, int(step)))

    def step_enqueue(self):
        capi.check(self._lib.mppi_step_enqueue(self._h))

    def step_wait(self):
        out = np.zeros(self.A, np.float32)
        capi.check(self._lib.mppi_step_wait(self._h, out.ctypes.data))
        return out

    def step_info(self):
        info = capi.MppiStepInfo()
        capi.check(self._lib.mppi_get_step_info(self._h, C.byref(info)))
        return dict(beta=np.float32(info.beta), eta=np.float32(info.eta),
                    argmin=int(info.argmin), step=int(info.step))

    def flags(self):
        """MPPI_FLAG_* bits the handle runs with (FLAG_AUTO_CHAIN resolved)."""
        f = C.c_uint32()
        capi.check(self._lib.mppi_get_flags(self._h, C.byref(f)))
        return int(f.value)

    def timer_start(self):
        capi.check(self._lib.mppi_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        capi.check(self._lib.mppi_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def set_profiling(self, on):
        capi.check(self._lib.mppi_set_profiling(self._h, int(bool(on))))

    def kernel_times(self):
        ms = (C.c_double * capi.K_COUNT)()
        n = (C.c_int64 * capi.K_COUNT)()
        capi.check(self._lib.mppi_get_kernel_times(self._h, ms, n))
        return {self._lib.mppi_kernel_name(i).decode(): (float(ms[i]), int(n[i]))
                for i in range(capi.K_COUNT)}

    def exchange_times(self):
        """MPPI_COMM_P2P: {push, wait, merge} microseconds of the last step's NVLink exchange."""
        us = (C.c_double * 3)()
        capi.check(self._lib.mppi_get_exchange_times(self._h, us))
        return {"push_us": float(us[0]), "wait_slowest_us": float(us[1]), "merge_us": float(us[2])}

    def launch_count(self):
        n = C.c_int64()
        capi.check(self._lib.mppi_get_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def p2p_handle(self) -> bytes:
        buf = (C.c_uint8 * capi.P2P_HANDLE_BYTES)()
        capi.check(self._lib.mppi_comm_p2p_handle(self._h, buf))
        return bytes(buf)

    def p2p_connect(self, handles: bytes):
        assert len(handles) == capi.P2P_HANDLE_BYTES * self.params.world_size
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        capi.check(self._lib.mppi_comm_p2p_connect(self._h, buf))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mppi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    lib = capi.load()
    buf = (C.c_uint8 * capi.COMM_ID_BYTES)()
    capi.check(lib.mppi_comm_unique_id(buf))
    return bytes(buf)
