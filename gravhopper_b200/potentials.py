"""Device-native analytic external potentials.

The reference sums Python callbacks every step (``Simulation.calculate_extra_acceleration``,
/root/reference/gravhopper/gravhopper.py:462-473); its docs use them for static background fields
(docs/source/examples.rst:74-75, 172-183: logarithmic halo, NFW).  A callback forces the step to
leave the GPU (x_half -> host -> Python -> device).  The objects below describe the common static
fields analytically; ``Simulation.add_external_force(obj)`` recognises them and hands them to the
engine (``gh_engine_add_potential``), which evaluates them at x_half on the device -- the run
stays fully device resident.

Every object is ALSO a valid reference-style hook: calling ``obj(pos, args)`` with an (N,3) length
Quantity (or kpc ndarray) returns the acceleration (Quantity in km/s/Myr, or ndarray), computed
on the host with the same formula.  Tests use that as the cross-check, and it is what the object
degrades to if it is registered together with Python callbacks in the reference itself.
"""
import numpy as np

from .units import u, has_units, to_value, LENUNIT, MASSUNIT, VELUNIT, ACCELUNIT

G = 4.30091727003628e-06                # kpc (km/s)^2 / Msun
KPC_PER_KMS_MYR = 1.022712165045695e-3  # 1 (km/s)^2/kpc = K km/s/Myr

POINTMASS, HERNQUIST, NFW_KIND, LOGHALO, MIYAMOTO = 1, 2, 3, 4, 5


class NativePotential(object):
    kind = 0

    def __init__(self, center=None):
        self.center = np.zeros(3) if center is None else np.asarray(to_value(center, LENUNIT), dtype=np.float64)

    def params(self):
        """The 8 doubles gh_engine_add_potential takes (kpc, Msun, km/s)."""
        raise NotImplementedError

    def _accel_kms2_per_kpc(self, d):
        raise NotImplementedError

    def acceleration(self, pos_kpc):
        """(N,3) kpc ndarray -> (N,3) km/s/Myr ndarray (host evaluation)."""
        d = np.asarray(pos_kpc, dtype=np.float64) - self.center
        return self._accel_kms2_per_kpc(d) * KPC_PER_KMS_MYR

    def __call__(self, pos, args=None):
        q = has_units(pos)
        a = self.acceleration(to_value(pos, LENUNIT))
        return u.Quantity(a, ACCELUNIT, copy=False) if q else a


class PointMass(NativePotential):
    """Point mass, optionally Plummer-softened: a = -G M d / (r^2 + b^2)^{3/2}."""
    kind = POINTMASS

    def __init__(self, mass, center=None, softening=0.0):
        NativePotential.__init__(self, center)
        self.mass = float(to_value(mass, MASSUNIT))
        self.b = float(to_value(softening, LENUNIT))

    def params(self):
        return [self.mass, self.center[0], self.center[1], self.center[2], self.b, 0, 0, 0]

    def _accel_kms2_per_kpc(self, d):
        s = (d ** 2).sum(axis=1) + self.b ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            w = np.where(s > 0, -G * self.mass / (s * np.sqrt(s)), 0.0)
        return d * w[:, None]


Plummer = PointMass


class Hernquist(NativePotential):
    """Hernquist sphere: a = -G M d / (r (r + a)^2)."""
    kind = HERNQUIST

    def __init__(self, mass, a, center=None):
        NativePotential.__init__(self, center)
        self.mass = float(to_value(mass, MASSUNIT))
        self.a = float(to_value(a, LENUNIT))

    def params(self):
        return [self.mass, self.center[0], self.center[1], self.center[2], self.a, 0, 0, 0]

    def _accel_kms2_per_kpc(self, d):
        r = np.sqrt((d ** 2).sum(axis=1))
        with np.errstate(divide="ignore", invalid="ignore"):
            w = np.where(r > 0, -G * self.mass / (r * (r + self.a) ** 2), 0.0)
        return d * w[:, None]


class NFW(NativePotential):
    """NFW halo given M_s = 4 pi rho0 rs^3 and rs: a = -G M_s (ln(1+x) - x/(1+x)) d / r^3, x = r/rs."""
    kind = NFW_KIND

    def __init__(self, mass_scale, rs, center=None):
        NativePotential.__init__(self, center)
        self.ms = float(to_value(mass_scale, MASSUNIT))
        self.rs = float(to_value(rs, LENUNIT))

    def params(self):
        return [self.ms, self.center[0], self.center[1], self.center[2], self.rs, 0, 0, 0]

    def _accel_kms2_per_kpc(self, d):
        r = np.sqrt((d ** 2).sum(axis=1))
        x = r / self.rs
        with np.errstate(divide="ignore", invalid="ignore"):
            w = np.where(r > 0, -G * self.ms * (np.log1p(x) - x / (1 + x)) / r ** 3, 0.0)
        return d * w[:, None]


class LogHalo(NativePotential):
    """Logarithmic halo Phi = v0^2/2 ln(rc^2 + x^2 + y^2 + z^2/q^2) (flat rotation curve v0)."""
    kind = LOGHALO

    def __init__(self, v0, rc=0.0, q=1.0, center=None):
        NativePotential.__init__(self, center)
        self.v0 = float(to_value(v0, VELUNIT))
        self.rc = float(to_value(rc, LENUNIT))
        self.q = float(q)

    def params(self):
        return [self.v0, self.center[0], self.center[1], self.center[2], self.rc, self.q, 0, 0]

    def _accel_kms2_per_kpc(self, d):
        iq2 = 1.0 / self.q ** 2
        w = -self.v0 ** 2 / (self.rc ** 2 + d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2 * iq2)
        return np.stack((w * d[:, 0], w * d[:, 1], w * d[:, 2] * iq2), axis=1)


class MiyamotoNagai(NativePotential):
    """Miyamoto-Nagai disk: Phi = -G M / sqrt(R^2 + (a + sqrt(z^2 + b^2))^2)."""
    kind = MIYAMOTO

    def __init__(self, mass, a, b, center=None):
        NativePotential.__init__(self, center)
        self.mass = float(to_value(mass, MASSUNIT))
        self.a = float(to_value(a, LENUNIT))
        self.b = float(to_value(b, LENUNIT))

    def params(self):
        return [self.mass, self.center[0], self.center[1], self.center[2], self.a, self.b, 0, 0]

    def _accel_kms2_per_kpc(self, d):
        zb = np.sqrt(d[:, 2] ** 2 + self.b ** 2)
        az = self.a + zb
        s = d[:, 0] ** 2 + d[:, 1] ** 2 + az ** 2
        w = -G * self.mass / (s * np.sqrt(s))
        with np.errstate(divide="ignore", invalid="ignore"):
            fz = np.where(zb > 0, w * d[:, 2] * az / zb, 0.0)
        return np.stack((w * d[:, 0], w * d[:, 1], fz), axis=1)
