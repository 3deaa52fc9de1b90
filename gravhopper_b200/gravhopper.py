"""Drop-in for the hot-path part of ``gravhopper.gravhopper``: the ``Simulation`` class (state,
leapfrog loop, acceleration dispatch, external-force hooks, ``add_IC``), the ``IC`` generators
and the exception types.  Reference: /root/reference/gravhopper/gravhopper.py:68-88 (exceptions),
:116-473 (Simulation core), :603-904 (hooks, add_IC), :1327-1785 (IC).

What is different underneath (and invisible through the API):

* ``run(N)`` does not loop in Python.  The state is uploaded once to a device-resident engine
  (libgravhopper_b200.so); each DKD step (gravhopper.py:405-416) is one fused force+kick+drift
  kernel sequence on the B200; snapshots stream back into ``positions``/``velocities`` while the
  following steps run.  Only when external-force callbacks are registered does the step split
  (half-drifted positions -> host -> callbacks -> device), because the callbacks are Python.
* Units: astropy if installed, else the stand-in in ``gravhopper_b200.units``; plain numbers are
  accepted and read as internal units (kpc, km/s, Msun, Myr).

Extra, optional constructor arguments (defaults reproduce the reference): ``precision``
('fp64'|'fp32'), ``theta`` (the reference hard-wires 0.7, gravhopper.py:446 + jbgrav.py:52),
``snapshot_every`` (the reference stores every step), ``device``.

Out of scope here (plotting, movies, pynbody/galpy/gala/agama adapters): the
corresponding methods raise ``ExternalPackageException``.
"""
import ctypes as C

import numpy as np

from . import _lib, ic_raw, jbgrav
from .potentials import NativePotential
from .units import (u, const, has_units, to_value, LENUNIT, VELUNIT, MASSUNIT, TIMEUNIT,
                    ACCELUNIT)

__all__ = ['Simulation', 'IC', 'GravHopperException', 'UninitializedSimulationException',
           'ICException', 'UnknownAlgorithmException', 'ExternalPackageException', 'force_centers']


class GravHopperException(Exception):
    """Parent class for all error exceptions."""
    pass


class UninitializedSimulationException(GravHopperException):
    """Exception for trying to run a simulation without any initial conditions."""
    pass


class ICException(GravHopperException):
    """Exception for trying to add ICs that don't make sense."""
    def __init__(self, msg):
        print(msg)


class UnknownAlgorithmException(GravHopperException):
    """Exception for using an unknown N-body algorithm name."""
    pass


class ExternalPackageException(GravHopperException):
    """Exception for trying to call a pynbody/galpy/gala/agama function when not using them."""
    def __init__(self, msg):
        print(msg)


def _q(arr, unit):
    """ndarray -> Quantity view in `unit` (no copy)."""
    return u.Quantity(arr, unit, copy=False)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class _Engine(object):
    """Thin owner of a gh_engine handle (single GPU)."""

    def __init__(self, n, precision, device=0):
        self.n = n
        self.precision = precision
        self.lib = _lib.lib()
        _lib.require_gpu()
        h = C.c_void_p()
        prec = _lib.GH_PREC_F64 if precision == 'fp64' else _lib.GH_PREC_F32
        _lib.check(self.lib.gh_engine_create(C.byref(h), device, n, 0, n, prec), "gh_engine_create")
        self.h = h

    def close(self):
        h, self.h = getattr(self, 'h', None), None
        if h:
            try:
                self.lib.gh_engine_destroy(h)
            except Exception:  # interpreter shutdown: the library may already be gone
                pass

    __del__ = close

    def set_potentials(self, pots):
        _lib.check(self.lib.gh_engine_clear_potentials(self.h))
        for p in pots:
            prm = (C.c_double * 8)(*[float(v) for v in p.params()])
            _lib.check(self.lib.gh_engine_add_potential(self.h, p.kind, prm, 8), "gh_engine_add_potential")

    def upload(self, pos, vel, mass):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        if self.precision == 'fp32':
            # fp32 coordinates are relative to an origin that starts at the mean position and moves
            # with the mean velocity
            origin = np.ascontiguousarray(pos.mean(axis=0))
            ovel = np.ascontiguousarray(vel.mean(axis=0))
            _lib.check(self.lib.gh_engine_set_origin(self.h, origin.ctypes.data_as(C.POINTER(C.c_double))))
            _lib.check(self.lib.gh_engine_set_origin_velocity(self.h, ovel.ctypes.data_as(C.POINTER(C.c_double))))
        _lib.check(self.lib.gh_engine_upload(self.h, _ptr(pos), _ptr(vel), _ptr(mass)), "gh_engine_upload")

    def run(self, nsteps, dt, eps, theta, alg, every, pos_hist, vel_hist):
        _lib.check(self.lib.gh_engine_run(self.h, nsteps, dt, eps, theta, alg, every,
                                          None if pos_hist is None else _ptr(pos_hist),
                                          None if vel_hist is None else _ptr(vel_hist)), "gh_engine_run")

    def prepare(self, dt):
        _lib.check(self.lib.gh_engine_prepare(self.h, dt), "gh_engine_prepare")

    def xhalf(self):
        out = np.empty((self.n, 3))
        _lib.check(self.lib.gh_engine_download_xhalf(self.h, _ptr(out)), "gh_engine_download_xhalf")
        return out

    def step(self, dt, eps, theta, alg, ext=None):
        if ext is not None:
            ext = np.ascontiguousarray(ext, dtype=np.float64)
        _lib.check(self.lib.gh_engine_step(self.h, dt, eps, theta, alg,
                                           None if ext is None else _ptr(ext), _lib.GH_MEM_HOST),
                   "gh_engine_step")

    def download(self, pos, vel):
        _lib.check(self.lib.gh_engine_download(self.h, _ptr(pos), _ptr(vel)), "gh_engine_download")

    def energy(self, eps):
        out = (C.c_double * 2)()
        _lib.check(self.lib.gh_engine_energy(self.h, eps, out), "gh_engine_energy")
        return out[0], out[1]

    def launches(self):
        n = C.c_int64()
        _lib.check(self.lib.gh_engine_launch_count(self.h, C.byref(n)))
        return n.value

    def last_force_ms(self):
        ms = C.c_float()
        _lib.check(self.lib.gh_engine_last_force_ms(self.h, C.byref(ms)))
        return ms.value


class _GroupEngine(object):
    """All GPUs of this process stepping one system (gh_group_create_local: one engine per device,
    ncclCommInitAll inside the library -- no launcher, no torch).  Targets are sharded evenly; for
    the tree the particles are first laid out in Morton blocks dealt round-robin over the devices
    (sharded.interleaved_layout) and mapped back at every snapshot."""

    def __init__(self, n, precision, devices):
        self.n, self.precision, self.devices = n, precision, list(devices)
        self.lib = _lib.lib()
        _lib.require_gpu()
        g = C.c_void_p()
        prec = _lib.GH_PREC_F64 if precision == 'fp64' else _lib.GH_PREC_F32
        devs = (C.c_int * len(self.devices))(*self.devices)
        _lib.check(self.lib.gh_group_create_local(C.byref(g), len(self.devices), devs, n, prec),
                   "gh_group_create_local")
        self.g = g
        self.parts = []
        for k in range(len(self.devices)):
            h, b, c = C.c_void_p(), C.c_int64(), C.c_int64()
            _lib.check(self.lib.gh_group_engine(g, k, C.byref(h), C.byref(b), C.byref(c)), "gh_group_engine")
            self.parts.append((h, b.value, c.value))
        self.perm = None

    def close(self):
        g, self.g = getattr(self, 'g', None), None
        if g:
            try:
                self.lib.gh_group_destroy(g)
            except Exception:  # interpreter shutdown
                pass

    __del__ = close

    def set_potentials(self, pots):
        for h, _, _ in self.parts:
            _lib.check(self.lib.gh_engine_clear_potentials(h))
            for p in pots:
                prm = (C.c_double * 8)(*[float(v) for v in p.params()])
                _lib.check(self.lib.gh_engine_add_potential(h, p.kind, prm, 8), "gh_engine_add_potential")

    def upload(self, pos, vel, mass, tree):
        from .sharded import interleaved_layout
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        self.perm = interleaved_layout(pos, len(self.parts)) if tree else None
        if self.perm is not None:
            pos, vel, mass = pos[self.perm], vel[self.perm], np.ascontiguousarray(mass[self.perm])
        origin = np.ascontiguousarray(pos.mean(axis=0))
        ovel = np.ascontiguousarray(vel.mean(axis=0))
        for h, b, c in self.parts:
            _lib.check(self.lib.gh_engine_set_origin(h, origin.ctypes.data_as(C.POINTER(C.c_double))))
            _lib.check(self.lib.gh_engine_set_origin_velocity(h, ovel.ctypes.data_as(C.POINTER(C.c_double))))
            p, w = np.ascontiguousarray(pos[b:b + c]), np.ascontiguousarray(vel[b:b + c])
            _lib.check(self.lib.gh_engine_upload(h, _ptr(p), _ptr(w), _ptr(mass)), "gh_engine_upload")

    def step(self, nsteps, dt, eps, theta, alg):
        _lib.check(self.lib.gh_group_step(self.g, int(nsteps), dt, eps, theta, alg), "gh_group_step")

    def download(self, pos_out, vel_out):
        """State of all devices into (N,3) rows, in the caller's particle order."""
        pos = pos_out if self.perm is None else np.empty_like(pos_out)
        vel = vel_out if self.perm is None else np.empty_like(vel_out)
        for h, b, c in self.parts:
            p, w = np.empty((c, 3)), np.empty((c, 3))
            _lib.check(self.lib.gh_engine_download(h, _ptr(p), _ptr(w)), "gh_engine_download")
            pos[b:b + c] = p
            vel[b:b + c] = w
        if self.perm is not None:
            pos_out[self.perm] = pos
            vel_out[self.perm] = vel


class Simulation(object):
    """Main class for N-body simulation (gravhopper.py:91-163).

    Attributes: Np, Nsnap, timestep, positions (Nsnap,Np,3), velocities (Nsnap,Np,3), masses (Np),
    times (Nsnap), lenunit, velunit, massunit, timeunit, accelunit, params.
    """

    def __init__(self, dt=1 * u.Myr, eps=100 * u.pc, algorithm='tree', precision='fp64', theta=0.7,
                 snapshot_every=1, device=0, devices=None, quadrupoles=False):
        self.ICarrays = False
        self.Np = 0
        self.Nsnap = 0
        self.timestep = 0
        self.running = False
        self._pos = None
        self._vel = None
        self._mass = None
        self._times = None
        self.extra_force_functions = []
        self.extra_timedependent_force_functions = []
        self.extra_velocitydependent_force_functions = []
        self.native_potentials = []  # analytic fields evaluated on the device (potentials.py)
        # Things can come in in various units, but use these internally (gravhopper.py:150-154)
        self.lenunit = LENUNIT
        self.velunit = VELUNIT
        self.massunit = MASSUNIT
        self.timeunit = TIMEUNIT
        self.accelunit = ACCELUNIT
        self.params = {}
        self.set_dt(dt)
        self.set_eps(eps)
        self.set_algorithm(algorithm)
        if precision not in ('fp64', 'fp32'):
            raise ValueError("precision must be 'fp64' or 'fp32'.")
        self.params['precision'] = precision
        self.params['theta'] = float(theta)
        if int(snapshot_every) < 1:
            raise ValueError("snapshot_every must be >= 1.")
        self.params['snapshot_every'] = int(snapshot_every)
        self.params['device'] = int(device)
        # quadrupoles: opt-in accuracy upgrade beyond the reference (tree only; see _jbgrav.tree_quadrupoles)
        self.params['quadrupoles'] = bool(quadrupoles)
        # devices: None / 1 = one GPU (`device`); an int N > 1 = GPUs 0..N-1; a list = those GPUs.
        # run() then shards the targets over them inside this process (no torchrun needed).
        if devices is None:
            devs = [int(device)]
        elif isinstance(devices, (list, tuple)):
            devs = [int(d) for d in devices]
        else:
            devs = list(range(int(devices)))
        if len(devs) < 1 or len(set(devs)) != len(devs):
            raise ValueError("devices must name at least one GPU, each once.")
        self.params['devices'] = devs
        self._engine = None
        self._group = None
        self._plot_parms = None

    # ---- history arrays as Quantity views (gravhopper.py:135-143) ---------------------------
    @property
    def positions(self):
        return None if self._pos is None else _q(self._pos, self.lenunit)

    @positions.setter
    def positions(self, value):
        self._pos = None if value is None else np.ascontiguousarray(to_value(value, self.lenunit))

    @property
    def velocities(self):
        return None if self._vel is None else _q(self._vel, self.velunit)

    @velocities.setter
    def velocities(self, value):
        self._vel = None if value is None else np.ascontiguousarray(to_value(value, self.velunit))

    @property
    def masses(self):
        return None if self._mass is None else _q(self._mass, self.massunit)

    @masses.setter
    def masses(self, value):
        self._mass = None if value is None else np.ascontiguousarray(to_value(value, self.massunit))

    @property
    def times(self):
        return None if self._times is None else _q(self._times, self.timeunit)

    @times.setter
    def times(self, value):
        self._times = None if value is None else np.ascontiguousarray(to_value(value, self.timeunit))

    # ---- parameters (gravhopper.py:166-290) -----------------------------------------------------
    def set_dt(self, dt):
        """Sets the simulation time step; ValueError if dt does not have dimensions of time."""
        if has_units(dt):
            try:
                _ = dt.to(u.Myr)
            except u.UnitConversionError:
                raise ValueError("dt must have dimensions of time.")
        self.params['dt'] = dt
        return

    def get_dt(self):
        return self.params['dt']

    def set_eps(self, eps):
        """Sets the softening length; ValueError if eps does not have dimensions of length."""
        if has_units(eps):
            try:
                _ = eps.to(u.kpc)
            except u.UnitConversionError:
                raise ValueError("eps must have dimensions of length.")
        self.params['eps'] = eps
        return

    def get_eps(self):
        return self.params['eps']

    def set_algorithm(self, algorithm):
        if algorithm in ('tree', 'direct'):
            self.params['algorithm'] = algorithm
        else:
            raise ValueError("algorithm must be 'tree' or 'direct'.")
        return

    def get_algorithm(self):
        return self.params['algorithm']

    def _dt_value(self):
        return float(to_value(self.params['dt'], self.timeunit))

    def _eps_value(self):
        return float(to_value(self.params['eps'], self.lenunit))

    def _alg_code(self):
        alg = self.params['algorithm']
        if alg == 'direct':
            return _lib.GH_ALG_DIRECT
        if alg == 'tree':
            return _lib.GH_ALG_TREE
        raise UnknownAlgorithmException()

    def _has_hooks(self):
        return bool(self.extra_force_functions or self.extra_timedependent_force_functions or
                    self.extra_velocitydependent_force_functions)

    def _get_engine(self):
        e = self._engine
        if e is None or e.n != self.Np or e.precision != self.params['precision']:
            if e is not None:
                e.close()
            e = _Engine(self.Np, self.params['precision'], self.params['device'])
            self._engine = e
        return e

    # ---- run loop (gravhopper.py:293-356) -------------------------------------------------------
    def run(self, N=1):
        """Run N timesteps.  Initializes a simulation that has not yet been run, or continues from
        the last snapshot if it has (gravhopper.py:293-320)."""
        if self.params.get('quadrupoles'):
            from . import _jbgrav
            was = _jbgrav.tree_quadrupoles()
            _jbgrav.tree_quadrupoles(True)   # process-wide switch of the library, restored afterwards
            try:
                return self._run(N)
            finally:
                _jbgrav.tree_quadrupoles(was)
        return self._run(N)

    def _run(self, N=1):
        N = int(N)
        every = self.params['snapshot_every']
        nnew = N // every + (1 if N % every else 0)
        if self.running == False:  # noqa: E712 (reference idiom)
            self.init_run(nnew)
        else:
            self.Nsnap += nnew
            self._pos = np.concatenate((self._pos, np.zeros((nnew, self.Np, 3))), axis=0)
            self._vel = np.concatenate((self._vel, np.zeros((nnew, self.Np, 3))), axis=0)
            self._times = np.concatenate((self._times, np.zeros((nnew))), axis=0)
        if N <= 0:
            return
        dt = self._dt_value()
        eps = self._eps_value()
        alg = self._alg_code()
        theta = self.params['theta']
        s0 = self.timestep
        if len(self.params['devices']) > 1 and not self._has_hooks() and self.Np >= len(self.params['devices']):
            self._run_group(N, nnew, every, dt, eps, theta, alg, s0)
            return
        eng = self._get_engine()
        eng.upload(self._pos[s0], self._vel[s0], self._mass)
        eng.set_potentials(self.native_potentials)
        if not self._has_hooks():
            # fully fused path: nothing leaves the device except the snapshots
            eng.run(N, dt, eps, theta, alg, every, self._pos[s0 + 1:s0 + 1 + nnew],
                    self._vel[s0 + 1:s0 + 1 + nnew])
            t0 = self._times[s0]
            for k in range(nnew):
                nsteps_done = min((k + 1) * every, N)
                # times[i] = times[i-1] + dt accumulated step by step (gravhopper.py:320)
                t = self._times[s0 + k]
                for _ in range(nsteps_done - k * every):
                    t = t + dt
                self._times[s0 + 1 + k] = t
            self.timestep = s0 + nnew
        else:
            t = self._times[s0]
            # the velocities of the previous step only travel to the host when a velocity-dependent
            # hook will read them (gravhopper.py:469-473); otherwise the state stays on the device
            # between snapshots and only x_half (the hooks' positions) is downloaded per step
            need_v = bool(self.extra_velocitydependent_force_functions)
            v_prev = self._vel[s0].copy() if need_v else None
            x_now = np.empty((self.Np, 3)) if need_v else None
            v_now = np.empty((self.Np, 3)) if need_v else None
            k = 0
            for step in range(1, N + 1):
                eng.prepare(dt)
                xh = eng.xhalf()
                kick_time = t + 0.5 * dt  # gravhopper.py:412
                ext = self._extra_accel_values(xh, kick_time, v_prev)
                eng.step(dt, eps, theta, alg, ext)
                t = t + dt
                if step % every == 0 or step == N:
                    eng.download(self._pos[s0 + 1 + k], self._vel[s0 + 1 + k])
                    if need_v:
                        v_prev = self._vel[s0 + 1 + k].copy()
                    self._times[s0 + 1 + k] = t
                    k += 1
                elif need_v:
                    eng.download(x_now, v_now)
                    v_prev = v_now.copy()
            self.timestep = s0 + nnew

    def _run_group(self, N, nnew, every, dt, eps, theta, alg, s0):
        """run() over several GPUs of this process (Simulation(devices=...)): the state stays on the
        devices between snapshots, each snapshot is one download of every device's slice."""
        g = self._group
        devs, prec = self.params['devices'], self.params['precision']
        if g is None or g.n != self.Np or g.precision != prec or g.devices != devs:
            if g is not None:
                g.close()
            g = _GroupEngine(self.Np, prec, devs)
            self._group = g
        g.upload(self._pos[s0], self._vel[s0], self._mass, tree=(alg == _lib.GH_ALG_TREE))
        g.set_potentials(self.native_potentials)
        done = 0
        for k in range(nnew):
            nst = min(every, N - done)
            g.step(nst, dt, eps, theta, alg)
            g.download(self._pos[s0 + 1 + k], self._vel[s0 + 1 + k])
            t = self._times[s0 + k]
            for _ in range(nst):
                t = t + dt   # times[i] = times[i-1] + dt accumulated step by step (gravhopper.py:320)
            self._times[s0 + 1 + k] = t
            done += nst
        self.timestep = s0 + nnew

    def init_run(self, Nsnap=None):
        """Initialize an N-body run (gravhopper.py:324-342)."""
        if self.ICarrays == False:  # noqa: E712
            raise UninitializedSimulationException
        self.Nsnap = Nsnap + 1
        self.Np = len(self.ICarrays['pos'])
        self._pos = np.zeros((self.Nsnap, self.Np, 3))
        self._vel = np.zeros((self.Nsnap, self.Np, 3))
        self._mass = np.zeros((self.Np))
        self._times = np.zeros((self.Nsnap))
        self._pos[0, :, :] = to_value(self.ICarrays['pos'], self.lenunit)
        self._vel[0, :, :] = to_value(self.ICarrays['vel'], self.velunit)
        self._mass[:] = to_value(self.ICarrays['mass'], self.massunit)
        self.running = True

    def reset(self):
        """Reset to the state before init_run(); ICs and external forces are preserved."""
        self.running = False
        self._pos = None
        self._vel = None
        self._mass = None
        self._times = None
        self.Nsnap = 0
        self.timestep = 0

    # ---- snapshots (gravhopper.py:361-402) ------------------------------------------------------
    def snap(self, step):
        return {'pos': self.positions[step, :, :], 'vel': self.velocities[step, :, :],
                'mass': self.masses[:]}

    def current_snap(self):
        return self.snap(self.timestep)

    def prev_snap(self):
        return self.snap(self.timestep - 1)

    # ---- one step / acceleration (gravhopper.py:405-473) ----------------------------------------
    def perform_timestep(self):
        """Advance by one snapshot with the DKD leapfrog (gravhopper.py:405-416).  Like the
        reference, expects ``timestep`` to have been incremented and the row to exist."""
        dt = self._dt_value()
        eng = self._get_engine()
        s = self.timestep
        eng.upload(self._pos[s - 1], self._vel[s - 1], self._mass)
        eng.set_potentials(self.native_potentials)
        eng.prepare(dt)
        ext = None
        if self._has_hooks():
            ext = self._extra_accel_values(eng.xhalf(), self._times[s - 1] + 0.5 * dt, self._vel[s - 1])
        eng.step(dt, self._eps_value(), self.params['theta'], self._alg_code(), ext)
        eng.download(self._pos[s], self._vel[s])

    def calculate_acceleration(self, time=None):
        """N-body acceleration at current_snap() plus external forces (gravhopper.py:419-459)."""
        prec = self.params['precision']
        if self.Np > 1:
            if self.params['algorithm'] == 'direct':
                nbody_gravity = jbgrav.direct_summation(self.current_snap(), self.params['eps'], precision=prec)
            elif self.params['algorithm'] == 'tree':
                nbody_gravity = jbgrav.tree_force(self.current_snap(), self.params['eps'],
                                                  theta=self.params['theta'], precision=prec)
            else:
                raise UnknownAlgorithmException()
        else:
            nbody_gravity = np.zeros((self.Np, 3)) * self.accelunit
        self._include_native_in_extra = True  # here the natives are evaluated on the host too
        try:
            extra_accel = self.calculate_extra_acceleration(self.current_snap()['pos'], nbody_gravity,
                                                            time=time, vel=self.prev_snap()['vel'])
        finally:
            self._include_native_in_extra = False
        return nbody_gravity + extra_accel

    def calculate_extra_acceleration(self, pos, template_array, time=None, vel=None):
        """Acceleration due to the registered external forces only (gravhopper.py:462-473)."""
        extaccel = np.zeros_like(template_array)
        if getattr(self, "_include_native_in_extra", False):
            for pot in self.native_potentials:
                extaccel += pot(pos, None)
        for fn, args in self.extra_force_functions:
            extaccel += fn(pos, args)
        for fn, args in self.extra_timedependent_force_functions:
            extaccel += fn(pos, time, args)
        for fn, args in self.extra_velocitydependent_force_functions:
            extaccel += fn(pos, vel, args)
        return extaccel

    def _extra_accel_values(self, xhalf, time_value, vel_values):
        """Call the hooks exactly as the reference does (Quantities in, acceleration Quantity out)
        and return the summed external acceleration in km/s/Myr as a plain array."""
        pos = _q(xhalf, self.lenunit)
        # vel_values is None when no velocity-dependent hook is registered (nothing reads it)
        vel = None if vel_values is None else _q(np.ascontiguousarray(vel_values), self.velunit)
        time = time_value * self.timeunit
        template = np.zeros((self.Np, 3)) * self.accelunit
        ext = self.calculate_extra_acceleration(pos, template, time=time, vel=vel)
        return to_value(ext, self.accelunit)

    # ---- hook registration (gravhopper.py:603-845) ----------------------------------------------
    def add_external_force(self, fn, args=None, agama_units=None):
        """Add an external position-dependent force ``fn(pos, args)`` (gravhopper.py:603-689).
        Lists are flattened like the reference does for composite potentials."""
        if isinstance(fn, list):
            for item in fn:
                self.add_external_force(item, args)
            return
        if isinstance(fn, NativePotential):
            if len(self.native_potentials) >= 4:
                self.extra_force_functions.append((fn, args))  # beyond 4: falls back to a callback
            else:
                self.native_potentials.append(fn)
            return
        if not callable(fn):
            raise ExternalPackageException("galpy/gala/agama potential objects are not supported "
                                           "by gravhopper_b200; wrap them in a function fn(pos, args).")
        self.extra_force_functions.append((fn, args))

    def add_external_timedependent_force(self, fn, agama_units=None, args=None):
        """Add an external force ``fn(pos, time, args)`` (gravhopper.py:693-786)."""
        if isinstance(fn, list):
            for item in fn:
                self.add_external_timedependent_force(item, args=args)
            return
        if not callable(fn):
            raise ExternalPackageException("galpy/gala/agama potential objects are not supported "
                                           "by gravhopper_b200; wrap them in a function.")
        self.extra_timedependent_force_functions.append((fn, args))

    def add_external_velocitydependent_force(self, fn, args=None):
        """Add an external force ``fn(pos, vel, args)`` (gravhopper.py:789-845)."""
        if isinstance(fn, list):
            for item in fn:
                self.add_external_velocitydependent_force(item, args)
            return
        if not callable(fn):
            raise ExternalPackageException("galpy dissipative-force objects are not supported by "
                                           "gravhopper_b200; wrap them in a function.")
        self.extra_velocitydependent_force_functions.append((fn, args))

    # ---- ICs (gravhopper.py:848-904) --------------------------------------------------------------
    def nrows(self, array):
        """Number of rows of an array that may be a single 3-vector (gravhopper.py:848-851)."""
        a = np.asarray(getattr(array, 'value', array))
        return 1 if a.ndim == 1 else a.shape[0]

    def add_IC(self, newIC):
        """Adds particles to the initial conditions (gravhopper.py:853-904)."""
        if 'pos' not in newIC:
            raise ICException("Missing 'pos' key in initial conditions function.")
        if 'vel' not in newIC:
            raise ICException("Missing 'vel' key in initial conditions function.")
        if 'mass' not in newIC:
            raise ICException("Missing 'mass' key in initial conditions function.")
        npos = self.nrows(newIC['pos'])
        nvel = self.nrows(newIC['vel'])
        nmass = len(newIC['mass'])
        if (npos != nvel) | (npos != nmass):
            raise ICException('Inconsistent number of particles in initial conditions function.')
        pos = np.atleast_2d(to_value(newIC['pos'], self.lenunit))
        vel = np.atleast_2d(to_value(newIC['vel'], self.velunit))
        mass = np.atleast_1d(to_value(newIC['mass'], self.massunit))
        new = {'pos': _q(pos, self.lenunit), 'vel': _q(vel, self.velunit),
               'mass': _q(mass, self.massunit)}
        if self.ICarrays == False:  # noqa: E712
            self.ICarrays = new
        else:
            self.ICarrays = {
                'pos': _q(np.vstack((to_value(self.ICarrays['pos'], self.lenunit), pos)), self.lenunit),
                'vel': _q(np.vstack((to_value(self.ICarrays['vel'], self.velunit), vel)), self.velunit),
                'mass': _q(np.hstack((to_value(self.ICarrays['mass'], self.massunit), mass)), self.massunit)}

    # ---- diagnostics (new; the reference has none) ------------------------------------------------
    def energy(self, step=None):
        """(KE, PE) of a stored snapshot in Msun (km/s)^2, with the potential consistent with the
        softened force law, computed on the GPU."""
        s = self.timestep if step is None else step
        eng = self._get_engine()
        eng.upload(self._pos[s], self._vel[s], self._mass)
        return eng.energy(self._eps_value())

    # ---- out of scope ---------------------------------------------------------------------------
    def pyn_snap(self, timestep=None):
        raise ExternalPackageException("pynbody conversion is outside gravhopper_b200's scope.")

    def plot_particles(self, *args, **kwargs):
        raise ExternalPackageException("plotting is outside gravhopper_b200's scope; "
                                       "use sim.positions with your own matplotlib code.")

    def movie_particles(self, *args, **kwargs):
        raise ExternalPackageException("movies are outside gravhopper_b200's scope.")


def force_centers(positions, velocities, center_pos=None, center_vel=None, force_origin=True):
    """Shift positions/velocities to the desired centre (gravhopper.py:1740-1785)."""
    lq, vq = has_units(positions), has_units(velocities)
    lu = positions.unit if lq else LENUNIT
    vu = velocities.unit if vq else VELUNIT
    cp = None if center_pos is None else to_value(center_pos, lu)
    cv = None if center_vel is None else to_value(center_vel, vu)
    p, v = ic_raw.force_centers(np.array(to_value(positions, lu)), np.array(to_value(velocities, vu)),
                                cp, cv, force_origin)
    return (_q(p, lu) if lq else p, _q(v, vu) if vq else v)


class IC(object):
    """Static functions that generate initial conditions for ``Simulation.add_IC``
    (gravhopper.py:1167-1734).  Arguments are Quantities (or plain numbers in kpc / Msun / km/s);
    the returned dict holds Quantities: 'pos' in the unit of the scale length given, 'vel' in
    km/s, 'mass' in the unit of the total mass given."""

    @staticmethod
    def _wrap(pos, vel, mass, len_like, mass_like):
        lu = len_like.unit if has_units(len_like) else LENUNIT
        mu = mass_like.unit if has_units(mass_like) else MASSUNIT
        return {'pos': _q(pos, LENUNIT).to(lu), 'vel': _q(vel, VELUNIT), 'mass': _q(mass, MASSUNIT).to(mu)}

    @staticmethod
    def _centers(center_pos, center_vel):
        cp = None if center_pos is None else to_value(center_pos, LENUNIT)
        cv = None if center_vel is None else to_value(center_vel, VELUNIT)
        return cp, cv

    @staticmethod
    def TSIS(N=None, maxrad=None, totmass=None, center_pos=None, center_vel=None, force_origin=True,
             seed=None):
        """Truncated singular isothermal sphere (gravhopper.py:1327-1400)."""
        if (N is None) or (maxrad is None) or (totmass is None):
            raise ICException("TSIS requires N, maxrad, and totmass.")
        cp, cv = IC._centers(center_pos, center_vel)
        p, v, m = ic_raw.TSIS(N, float(to_value(maxrad, LENUNIT)), float(to_value(totmass, MASSUNIT)),
                              cp, cv, force_origin, seed)
        return IC._wrap(p, v, m, maxrad, totmass)

    @staticmethod
    def Plummer(N=None, b=None, totmass=None, center_pos=None, center_vel=None, force_origin=True,
                seed=None):
        """Isotropic Plummer model (gravhopper.py:1404-1493)."""
        if (N is None) or (b is None) or (totmass is None):
            raise ICException("Plummer requires N, b, and totmass.")
        cp, cv = IC._centers(center_pos, center_vel)
        p, v, m = ic_raw.Plummer(N, float(to_value(b, LENUNIT)), float(to_value(totmass, MASSUNIT)),
                                 cp, cv, force_origin, seed)
        return IC._wrap(p, v, m, b, totmass)

    @staticmethod
    def Hernquist(N=None, a=None, totmass=None, cutoff=10., center_pos=None, center_vel=None,
                  force_origin=True, seed=None):
        """Isotropic Hernquist model (gravhopper.py:1497-1607)."""
        if (N is None) or (a is None) or (totmass is None):
            raise ICException("Hernquist requires N, a, and totmass.")
        cp, cv = IC._centers(center_pos, center_vel)
        p, v, m = ic_raw.Hernquist(N, float(to_value(a, LENUNIT)), float(to_value(totmass, MASSUNIT)),
                                   cutoff, cp, cv, force_origin, seed)
        return IC._wrap(p, v, m, a, totmass)

    @staticmethod
    def expdisk(sigma0=None, Rd=None, z0=None, sigmaR_Rd=None, external_rotcurve=None, N=None,
                center_pos=None, center_vel=None, force_origin=True, seed=None):
        """Exponential disk (gravhopper.py:1611-1734).  ``external_rotcurve`` takes a length
        Quantity (or kpc) and returns a velocity Quantity (or km/s), as in the reference."""
        if (N is None) or (sigma0 is None) or (Rd is None) or (z0 is None) or (sigmaR_Rd is None):
            raise ICException("expdisk requires N, sigma0, Rd, z0, and sigmaR_Rd.")
        cp, cv = IC._centers(center_pos, center_vel)
        rot = None
        if external_rotcurve is not None:
            def rot(R_kpc):
                return to_value(external_rotcurve(_q(np.asarray(R_kpc, dtype=np.float64), LENUNIT)), VELUNIT)
        s0 = float(to_value(sigma0, MASSUNIT / LENUNIT ** 2))
        p, v, m = ic_raw.expdisk(N, s0, float(to_value(Rd, LENUNIT)), float(to_value(z0, LENUNIT)),
                                 float(to_value(sigmaR_Rd, VELUNIT)), rot, cp, cv, force_origin, seed)
        return IC._wrap(p, v, m, Rd, None)

    @staticmethod
    def from_galpy_df(*args, **kwargs):
        raise ExternalPackageException("galpy is outside gravhopper_b200's scope.")

    @staticmethod
    def from_pyn_snap(*args, **kwargs):
        raise ExternalPackageException("pynbody is outside gravhopper_b200's scope.")
