"""Units for the drop-in API: astropy when it is installed, otherwise a small built-in stand-in.

The reference's public API speaks astropy Quantities (``/root/reference/gravhopper/jbgrav.py:38-48``,
``gravhopper.py:150-154``).  astropy is not installed on the B200 image, so ``u`` and ``const``
here are either astropy's own modules or a minimal implementation covering what GravHopper's API
needs: the length/mass/time units of its docs, ``Quantity`` arithmetic with conversion,
``.to()``, ``.value``, ``.unit``, ``UnitConversionError`` and ``const.G``.  The stand-in uses
astropy's definitions (IAU 2012/2015, CODATA 2018) so that conversion factors agree with astropy's
to the last bit that matters (see tests/test_units.py: C_ACC and KPC_PER_KMS_MYR).

Internally the engine works on plain float64 arrays in kpc, km/s, Msun, Myr; units only appear at
the API boundary.
"""
import numpy as np

try:  # pragma: no cover - exercised only where astropy exists
    from astropy import units as u  # type: ignore
    from astropy import constants as const  # type: ignore
    HAVE_ASTROPY = True
except ImportError:
    HAVE_ASTROPY = False

if not HAVE_ASTROPY:
    import types

    class UnitConversionError(ValueError):
        pass

    class Unit(object):
        """scale to SI and integer/fractional exponents of (length, mass, time)."""
        __array_priority__ = 20000

        def __init__(self, scale, dims, name=None):
            self.scale = float(scale)
            self.dims = tuple(dims)
            self.name = name

        def __repr__(self):
            if self.name:
                return self.name
            return "Unit(%r, %r)" % (self.scale, self.dims)
        __str__ = __repr__

        def _compose(self, other, sign):
            return Unit(self.scale * other.scale ** sign,
                        tuple(a + sign * b for a, b in zip(self.dims, other.dims)),
                        _compose_name(self, other, sign))

        def __mul__(self, other):
            if isinstance(other, Unit):
                return self._compose(other, 1)
            if isinstance(other, Quantity):
                return Quantity(other.value, self * other.unit)
            return Quantity(other, self)

        def __rmul__(self, other):
            if isinstance(other, Quantity):
                return Quantity(other.value, other.unit * self)
            return Quantity(other, self)

        def __truediv__(self, other):
            if isinstance(other, Unit):
                return self._compose(other, -1)
            if isinstance(other, Quantity):
                return Quantity(1.0 / other.value, self / other.unit)
            return Quantity(1.0 / np.asarray(other, dtype=np.float64), self)

        def __rtruediv__(self, other):
            inv = Unit(1.0 / self.scale, tuple(-d for d in self.dims),
                       "1 / (%s)" % self.name if self.name else None)
            if isinstance(other, Quantity):
                return Quantity(other.value, other.unit * inv)
            return Quantity(other, inv)

        def __pow__(self, p):
            return Unit(self.scale ** p, tuple(d * p for d in self.dims),
                        "(%s)**%s" % (self.name, p) if self.name else None)

        def __eq__(self, other):
            return isinstance(other, Unit) and self.dims == other.dims and \
                np.isclose(self.scale, other.scale, rtol=1e-14, atol=0)

        def __hash__(self):
            return hash(self.dims)

        def is_equivalent(self, other):
            other = getattr(other, "unit", other)
            return self.dims == other.dims

        def to(self, other, value=1.0):
            other = getattr(other, "unit", other)
            if self.dims != other.dims:
                raise UnitConversionError("'%s' and '%s' are not convertible" % (self, other))
            return value * _snap(self.scale / other.scale)

        @property
        def unit(self):
            return self

    def _snap(f):
        """Round a scale ratio that is within 2 ulp of a short decimal (0.001, 1e6, 3.6 ...) to it,
        as astropy's prefix arithmetic yields; leave genuinely irrational ratios alone."""
        g = float("%.15g" % f)
        return g if abs(g - f) <= 4.5e-16 * abs(f) else f

    def _compose_name(a, b, sign):
        if a.name is None or b.name is None:
            return None
        if a.name == "":
            return b.name if sign > 0 else "1 / %s" % b.name
        return "%s %s" % (a.name, b.name) if sign > 0 else "%s / %s" % (a.name, b.name)

    def _mkarr(value, dtype, copy):
        if copy:
            return np.array(value, dtype=dtype or np.float64)
        return np.asarray(value, dtype=dtype or np.float64)

    def _unit_of(x):
        return x.unit if isinstance(x, Quantity) else (x if isinstance(x, Unit) else dimensionless_unscaled)

    def _val(x):
        if isinstance(x, Quantity):
            return x.view(np.ndarray)
        return x

    class Quantity(np.ndarray):
        __array_priority__ = 10000

        def __new__(cls, value, unit=None, dtype=None, copy=True):
            if isinstance(value, Quantity):
                if unit is None:
                    unit = value.unit
                    arr = _mkarr(value.view(np.ndarray), dtype, copy)
                else:
                    arr = np.array(value.to(unit).view(np.ndarray), dtype=dtype or np.float64)
            elif isinstance(value, (list, tuple)) and any(isinstance(v, Quantity) for v in value):
                unit = unit or value[0].unit
                arr = np.array([Quantity(v, unit).view(np.ndarray) for v in value], dtype=np.float64)
            else:
                arr = _mkarr(value, dtype, copy)
            obj = arr.view(cls)
            obj._unit = unit if unit is not None else dimensionless_unscaled
            return obj

        def __array_finalize__(self, obj):
            self._unit = getattr(obj, "_unit", None) or dimensionless_unscaled

        @property
        def unit(self):
            return self._unit

        @property
        def value(self):
            v = self.view(np.ndarray)
            return v if v.ndim else v[()]

        def to(self, unit):
            unit = getattr(unit, "unit", unit)
            f = self._unit.to(unit)
            return Quantity(self.view(np.ndarray) * f, unit)

        def to_value(self, unit):
            return self.to(unit).value

        @property
        def si(self):
            return Quantity(self.view(np.ndarray) * self._unit.scale, Unit(1.0, self._unit.dims))

        def decompose(self):
            return self.si

        def __repr__(self):
            return "<Quantity %s %s>" % (np.ndarray.__repr__(self.view(np.ndarray)), self._unit)

        def __str__(self):
            return "%s %s" % (self.view(np.ndarray), self._unit)

        def __format__(self, spec):
            if self.ndim == 0:
                return format(float(self.view(np.ndarray)), spec) + " %s" % self._unit
            return str(self)

        def __getitem__(self, item):
            out = np.ndarray.__getitem__(self.view(np.ndarray), item)
            return Quantity(out, self._unit, copy=False) if isinstance(out, np.ndarray) \
                else Quantity(out, self._unit)

        def __setitem__(self, item, value):
            if isinstance(value, Quantity):
                value = value.to(self._unit).view(np.ndarray)
            elif not self._unit.dims == (0, 0, 0) and not np.all(np.asarray(value) == 0):
                raise UnitConversionError("cannot assign a dimensionless value to a Quantity")
            np.ndarray.__setitem__(self.view(np.ndarray), item, value)

        def __iter__(self):
            for k in range(len(self)):
                yield self[k]

        def __reduce__(self):
            return (Quantity, (self.view(np.ndarray).copy(), self._unit))

        _SAME_UNIT = ("add", "subtract", "maximum", "minimum", "fmax", "fmin", "hypot", "remainder")
        _COMPARE = ("less", "less_equal", "greater", "greater_equal", "equal", "not_equal")
        _KEEP = ("negative", "positive", "absolute", "fabs", "rint", "floor", "ceil", "trunc",
                 "conjugate")
        _DIMLESS = ("exp", "log", "log10", "log2", "sin", "cos", "tan", "arcsin", "arccos", "arctan",
                    "arctanh", "tanh", "sinh", "cosh", "expm1", "log1p", "arctan2")

        def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
            name = ufunc.__name__
            out = kwargs.pop("out", None)
            units = [_unit_of(x) for x in inputs]
            vals = [_val(x) for x in inputs]
            if method == "reduce":
                res = getattr(ufunc, method)(vals[0], **kwargs)
                ru = units[0] if name in ("add", "maximum", "minimum") else None
                return self._wrap(res, ru, out)
            if method != "__call__":
                return NotImplemented
            if name in self._SAME_UNIT or name in self._COMPARE:
                base = units[0]
                if not isinstance(inputs[0], Quantity) and np.all(np.asarray(vals[0]) == 0):
                    base = units[1]
                conv = []
                for x, un, v in zip(inputs, units, vals):
                    if un.dims != base.dims:
                        if not isinstance(x, Quantity) and np.all(np.asarray(v) == 0):
                            conv.append(v)
                            continue
                        raise UnitConversionError("'%s' and '%s' are not convertible" % (un, base))
                    f = _snap(un.scale / base.scale)
                    conv.append(v if f == 1.0 else np.asarray(v) * f)
                res = ufunc(*conv, **kwargs)
                return self._wrap(res, None if name in self._COMPARE else base, out)
            if name == "multiply":
                return self._wrap(ufunc(*vals, **kwargs), units[0] * units[1], out)
            if name in ("true_divide", "divide"):
                return self._wrap(ufunc(*vals, **kwargs), units[0] / units[1], out)
            if name == "sqrt":
                return self._wrap(ufunc(*vals, **kwargs), units[0] ** 0.5, out)
            if name == "square":
                return self._wrap(ufunc(*vals, **kwargs), units[0] ** 2, out)
            if name == "reciprocal":
                return self._wrap(ufunc(*vals, **kwargs), units[0] ** -1, out)
            if name == "power":
                p = vals[1]
                if np.ndim(p) != 0 and units[0].dims != (0, 0, 0):
                    raise ValueError("Quantity ** array needs a dimensionless base")
                pu = units[0] ** float(p) if np.ndim(p) == 0 else units[0]
                return self._wrap(ufunc(*vals, **kwargs), pu, out)
            if name in self._KEEP or name in ("isfinite", "isnan", "isinf", "sign", "signbit"):
                keep = units[0] if name in self._KEEP else None
                return self._wrap(ufunc(*vals, **kwargs), keep, out)
            if name in self._DIMLESS:
                conv = []
                for un, v in zip(units, vals):
                    if un.dims != (0, 0, 0) and name != "arctan2":
                        raise UnitConversionError("%s needs a dimensionless argument" % name)
                    conv.append(np.asarray(v) * un.scale if un.dims == (0, 0, 0) and un.scale != 1.0 else v)
                return self._wrap(ufunc(*conv, **kwargs), None, out)
            return NotImplemented

        @staticmethod
        def _wrap(res, unit, out):
            if out is not None:
                o = out[0]
                if isinstance(o, Quantity):
                    if unit is not None and unit.dims != o._unit.dims:
                        raise UnitConversionError("'%s' and '%s' are not convertible" % (unit, o._unit))
                    f = 1.0 if unit is None else _snap(unit.scale / o._unit.scale)
                    np.copyto(o.view(np.ndarray), np.asarray(res) * f)
                else:
                    np.copyto(o, res)
                return o
            if unit is None:
                return res
            if unit.dims == (0, 0, 0) and unit.scale == 1.0 and False:
                return res
            return Quantity(res, unit, copy=False)

        def __array_function__(self, func, types, args, kwargs):
            name = func.__name__
            if name in ("concatenate", "vstack", "hstack", "stack", "append"):
                seq = args[0] if name != "append" else list(args[:2])
                base = None
                for s in seq:
                    if isinstance(s, Quantity):
                        base = s.unit
                        break
                conv = [Quantity(s, base).view(np.ndarray) if isinstance(s, Quantity)
                        else np.asarray(s) for s in seq]
                if name == "append":
                    return Quantity(func(conv[0], conv[1], *args[2:], **kwargs), base, copy=False)
                return Quantity(func(conv, *args[1:], **kwargs), base, copy=False)
            if name in ("mean", "sum", "median", "std", "amax", "amin", "max", "min", "cumsum",
                        "sort", "ravel", "transpose", "reshape", "squeeze", "atleast_1d",
                        "atleast_2d", "copy", "nanmean", "nansum", "diff", "average", "broadcast_to",
                        "expand_dims", "flip", "roll", "take", "tile", "repeat"):
                a = args[0]
                res = func(a.view(np.ndarray), *args[1:], **kwargs)
                return Quantity(res, a.unit, copy=False)
            if name in ("var",):
                a = args[0]
                return Quantity(func(a.view(np.ndarray), *args[1:], **kwargs), a.unit ** 2)
            if name in ("zeros_like", "ones_like", "empty_like", "full_like"):
                a = args[0]
                return Quantity(func(a.view(np.ndarray), *args[1:], **kwargs), a.unit, copy=False)
            if name in ("shape", "size", "ndim", "argsort", "argmax", "argmin", "nonzero",
                        "isclose", "allclose", "array_equal", "any", "all", "count_nonzero",
                        "iscomplexobj", "isrealobj", "result_type", "can_cast"):
                conv = [a.view(np.ndarray) if isinstance(a, Quantity) else a for a in args]
                if name in ("isclose", "allclose", "array_equal") and isinstance(args[1], Quantity) \
                        and isinstance(args[0], Quantity):
                    conv[1] = args[1].to(args[0].unit).view(np.ndarray)
                return func(*conv, **kwargs)
            if name == "norm":
                a = args[0]
                return Quantity(func(a.view(np.ndarray), *args[1:], **kwargs), a.unit, copy=False)
            if name == "where":
                cond = np.asarray(args[0])
                if len(args) == 1:
                    return func(cond)
                base = _unit_of(args[1])
                x = _val(args[1])
                y = Quantity(args[2], None).to(base).view(np.ndarray) if isinstance(args[2], Quantity) \
                    else args[2]
                return Quantity(func(cond, x, y), base, copy=False)
            if name in ("dot", "cross", "matmul", "outer", "inner"):
                ua, ub = _unit_of(args[0]), _unit_of(args[1])
                return Quantity(func(_val(args[0]), _val(args[1]), *args[2:], **kwargs), ua * ub)
            conv = [a.view(np.ndarray) if isinstance(a, Quantity) else a for a in args]
            return func(*conv, **kwargs)

    dimensionless_unscaled = Unit(1.0, (0, 0, 0), "")

    def _mk(scale, dims, name):
        return Unit(scale, dims, name)

    _L, _M, _T = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    u = types.ModuleType("gravhopper_b200.units.u")
    u.Unit = Unit
    u.Quantity = Quantity
    u.UnitConversionError = UnitConversionError
    u.dimensionless_unscaled = dimensionless_unscaled
    u.one = dimensionless_unscaled
    u.m = _mk(1.0, _L, "m")
    u.cm = _mk(1e-2, _L, "cm")
    u.km = _mk(1e3, _L, "km")
    u.au = u.AU = _mk(1.495978707e11, _L, "AU")
    u.pc = _mk(3.0856775814913674e16, _L, "pc")
    u.kpc = _mk(3.0856775814913674e19, _L, "kpc")
    u.Mpc = _mk(3.0856775814913674e22, _L, "Mpc")
    u.lyr = _mk(9.4607304725808e15, _L, "lyr")
    u.Rsun = u.R_sun = _mk(6.957e8, _L, "Rsun")
    u.kg = _mk(1.0, _M, "kg")
    u.g = _mk(1e-3, _M, "g")
    u.Msun = u.M_sun = u.solMass = _mk(1.988409870698051e30, _M, "Msun")
    u.Mearth = u.M_earth = _mk(5.972167867791379e24, _M, "Mearth")
    u.Mjup = u.M_jup = _mk(1.8981245973360505e27, _M, "Mjup")
    u.s = _mk(1.0, _T, "s")
    u.min = _mk(60.0, _T, "min")
    u.h = u.hr = u.hour = _mk(3600.0, _T, "h")
    u.d = u.day = _mk(86400.0, _T, "d")
    u.yr = u.year = _mk(31557600.0, _T, "yr")
    u.kyr = _mk(31557600.0e3, _T, "kyr")
    u.Myr = _mk(31557600.0e6, _T, "Myr")
    u.Gyr = _mk(31557600.0e9, _T, "Gyr")
    u.rad = _mk(1.0, (0, 0, 0), "rad")
    u.deg = _mk(np.pi / 180.0, (0, 0, 0), "deg")

    const = types.ModuleType("gravhopper_b200.units.const")
    const.G = Quantity(6.6743e-11, u.m ** 3 / (u.kg * u.s ** 2))
    const.c = Quantity(299792458.0, u.m / u.s)
    const.M_sun = Quantity(1.0, u.Msun)
    const.au = Quantity(1.0, u.au)
    const.pc = Quantity(1.0, u.pc)


def has_units(x):
    """True for astropy (or stand-in) Quantities."""
    return hasattr(x, "unit") and hasattr(x, "to") and hasattr(x, "value")


def to_value(x, unit):
    """x in `unit` as a float64 ndarray / float.  Plain numbers are taken to be in `unit` already
    (lets the package be used without any units machinery)."""
    if has_units(x):
        return np.asarray(x.to(unit).value, dtype=np.float64)
    return np.asarray(x, dtype=np.float64)


# GravHopper's internal units (gravhopper.py:150-154)
LENUNIT = u.kpc
VELUNIT = u.km / u.s
MASSUNIT = u.Msun
TIMEUNIT = u.Myr
ACCELUNIT = VELUNIT / TIMEUNIT
