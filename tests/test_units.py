"""CPU tests: the units layer (astropy or the stand-in) yields the constants the engine uses."""
import numpy as np
import pytest

from gravhopper_b200 import jbgrav
from gravhopper_b200.units import u, const, HAVE_ASTROPY


def test_hot_path_unit_factors():
    # jbgrav.py:40,48 and gravhopper.py:409 (values from astropy's definitions)
    assert jbgrav.C_ACC == 4.398600412921223e-09
    f = (const.G * u.Msun / u.kpc ** 2).to(u.km / u.s / u.Myr).value
    assert abs(f - 4.398600412921223e-09) <= 1e-15 * f
    k = (1 * u.km / u.s * u.Myr).to(u.kpc).value
    assert abs(k - 1.022712165045695e-3) <= 1e-15 * k
    g = const.G.to(u.kpc * (u.km / u.s) ** 2 / u.Msun).value
    assert abs(g - 4.30091727003628e-06) <= 1e-15 * g


def test_quantity_behaviour_used_by_the_api():
    x = np.array([[1, 2, 3.], [4, 5, 6]]) * u.pc
    assert x.to(u.kpc).value[0, 0] == 0.001
    assert (100 * u.pc).to(u.kpc).value == 0.1
    assert (5e3 * u.yr).to(u.Myr).value == 0.005
    y = np.zeros((2, 3)) * u.kpc
    y[:] = x  # assignment converts (gravhopper.py:338)
    assert np.allclose(y.value, x.value * 1e-3)
    assert np.vstack((x, y)).shape == (4, 3)
    with pytest.raises(u.UnitConversionError):
        x.to(u.s)
    r = np.sqrt((x ** 2).sum(axis=1))
    assert np.allclose(r.to(u.pc).value, np.sqrt((x.value ** 2).sum(axis=1)))
    a = (const.G * (1e8 * u.Msun) / (10 * u.kpc) ** 2).to(u.km / u.s / u.Myr)
    assert np.isclose(a.value, 4.398600412921223e-09 * 1e8 / 100)


@pytest.mark.skipif(HAVE_ASTROPY, reason="stand-in only")
def test_standin_rejects_mixed_dimensions():
    with pytest.raises(u.UnitConversionError):
        _ = (1 * u.kpc) + (1 * u.s)
