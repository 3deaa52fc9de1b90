"""Drop-in for ``gravhopper.jbgrav`` (/root/reference/gravhopper/jbgrav.py:14-177): the four
public force functions on snapshot dicts of Quantities.  Same names, same argument order and
defaults (``theta=0.7`` lives here, jbgrav.py:52), same units in and out:

    in : snap['pos'] (Np,3) length, snap['mass'] (Np) mass, eps length, force_pos (N,3) length
    out: (Np|N, 3) acceleration Quantity in km/s/Myr  (jbgrav.py:40-48)

The unit strip and the final scale factor G*Msun/kpc^2 -> km/s/Myr follow jbgrav.py:38-48; the
force itself runs on the B200 through ``gravhopper_b200._jbgrav``.  Without astropy the
built-in stand-in Quantity (``gravhopper_b200.units``) is used; plain ndarrays are accepted too
and taken to be in kpc / Msun, in which case a plain ndarray in km/s/Myr comes back.
"""
import numpy as np

from . import _jbgrav
from .units import u, const, has_units, to_value

__all__ = ['direct_summation', 'direct_summation_position', 'tree_force', 'tree_force_position']

# jbgrav.py:38-41
_unit_length = u.kpc
_unit_mass = u.Msun
_desired_accel_unit = u.km / u.s / u.Myr


from .units import HAVE_ASTROPY  # noqa: E402


def _accel_factor():
    if HAVE_ASTROPY:
        unit_accel = const.G * _unit_mass / (_unit_length ** 2)
        return unit_accel.to(_desired_accel_unit)  # Quantity, value 4.398600412921223e-09
    # the stand-in's product of scales rounds one ulp differently; pin astropy's value
    return u.Quantity(4.398600412921223e-09, _desired_accel_unit)


_ACCEL_FACTOR = _accel_factor()
C_ACC = float(_ACCEL_FACTOR.value)


def _strip(snap, eps, force_pos=None):
    positions = snap['pos']
    masses = snap['mass']
    quant = has_units(positions) or has_units(masses) or has_units(eps)
    posarray = to_value(positions, _unit_length)
    massarray = to_value(masses, _unit_mass)
    eps_in_units = float(to_value(eps, _unit_length))
    fp = None if force_pos is None else to_value(force_pos, _unit_length)
    return posarray, massarray, eps_in_units, fp, quant


def _finish(forcearray, quant):
    if quant:
        return forcearray * _ACCEL_FACTOR  # jbgrav.py:48
    return forcearray * C_ACC


def direct_summation(snap, eps, precision='fp64'):
    """Gravitational acceleration on every particle from every other particle by direct
    summation (jbgrav.py:14-48)."""
    pos, mass, e, _, q = _strip(snap, eps)
    return _finish(_jbgrav.direct_summation(pos, mass, e, precision=precision), q)


def tree_force(snap, eps, theta=0.7, precision='fp64'):
    """Same with a Barnes-Hut tree, opening angle ``theta`` (jbgrav.py:52-88)."""
    pos, mass, e, _, q = _strip(snap, eps)
    return _finish(_jbgrav.tree_force(pos, mass, e, theta, precision=precision), q)


def direct_summation_position(snap, force_pos, eps, precision='fp64'):
    """Acceleration at ``force_pos`` from every particle in the snapshot (jbgrav.py:91-133)."""
    pos, mass, e, fp, q = _strip(snap, eps, force_pos)
    q = q or has_units(force_pos)
    return _finish(_jbgrav.direct_summation_position(pos, mass, fp, e, precision=precision), q)


def tree_force_position(snap, force_pos, eps, theta=0.7, precision='fp64'):
    """Tree acceleration at ``force_pos`` (jbgrav.py:136-177)."""
    pos, mass, e, fp, q = _strip(snap, eps, force_pos)
    q = q or has_units(force_pos)
    return _finish(_jbgrav.tree_force_position(pos, mass, fp, e, theta, precision=precision), q)
