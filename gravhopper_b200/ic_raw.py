"""Host-side initial-condition samplers in GravHopper's internal units (kpc, km/s, Msun).

These are the synthetic inputs of the hot path (SURVEY 8a "inputs", 8d).  They implement the
same sampling maths and the same ``np.random.default_rng(seed)`` draw order as the reference's
``IC`` static methods, without astropy, and without the three defects that stop the reference
generators from running at the benchmark sizes:

* ``Plummer``   -- /root/reference/gravhopper/gravhopper.py:1452-1491
* ``Hernquist`` -- :1544-1605 (the discarded ``np.append`` at :1573-1575 is really applied)
* ``expdisk``   -- :1669-1733 (radial CDF table extended to R=0; central-difference derivative
  instead of the removed ``scipy.misc.derivative`` signature)
* ``TSIS``      -- :1378-1398
* ``force_centers`` -- :1768-1785

All functions return ``(pos (N,3) kpc, vel (N,3) km/s, mass (N,) Msun)`` as float64 ndarrays.
The unit-aware ``IC`` class in ``gravhopper_b200/gravhopper.py`` wraps them.
"""
import numpy as np

# G in kpc (km/s)^2 / Msun  (astropy const.G with CODATA 2018)
G = 4.30091727003628e-06


def _interp_monotone(xp, fp):
    """interp1d(xp, fp) for monotonically non-decreasing xp (linear, bounds clipped)."""
    xp = np.asarray(xp, dtype=np.float64)
    fp = np.asarray(fp, dtype=np.float64)

    def f(x):
        return np.interp(x, xp, fp)
    return f


def force_centers(pos, vel, center_pos=None, center_vel=None, force_origin=True):
    """gravhopper.py:1768-1785: shift to the requested (unweighted) mean position/velocity."""
    if force_origin:
        if center_pos is None:
            center_pos = np.zeros(3)
        if center_vel is None:
            center_vel = np.zeros(3)
    if center_pos is not None:
        pos += np.asarray(center_pos, dtype=np.float64) - np.mean(pos, axis=0)
    if center_vel is not None:
        vel += np.asarray(center_vel, dtype=np.float64) - np.mean(vel, axis=0)
    return pos, vel


def _sphere(rng, radius, N):
    costheta = rng.uniform(-1.0, 1.0, size=N)
    phi = rng.uniform(0.0, 2.0 * np.pi, size=N)
    sintheta = np.sqrt(1.0 - costheta ** 2)
    return np.vstack((radius * sintheta * np.cos(phi), radius * sintheta * np.sin(phi),
                      radius * costheta)).T


def Plummer(N, b, totmass, center_pos=None, center_vel=None, force_origin=True, seed=None):
    """Isotropic Plummer sphere; b in kpc, totmass in Msun (gravhopper.py:1452-1491)."""
    rng = np.random.default_rng(seed)
    rad_xi = rng.uniform(0.0, 1.0, size=N)
    radius = b / np.sqrt(rad_xi ** (-2. / 3) - 1)
    pos = _sphere(rng, radius, N)
    # Aarseth+ 1974: q from q^2 (1-q^2)^(7/2) by a 101-point tabulated inverse CDF
    qax = np.arange(0, 1.01, 0.01)
    q_prob = qax ** 2 * (1. - qax ** 2) ** (3.5)
    q_cumprob = np.cumsum(q_prob)
    q_cumprob /= q_cumprob[-1]
    vel_xi = rng.uniform(0.0, 1.0, size=N)
    q = _interp_monotone(q_cumprob, qax)(vel_xi)
    velocity = q * np.sqrt(2. * G * totmass / b) * (1. + (radius / b) ** 2) ** (-0.25)
    vel = _sphere(rng, velocity, N)
    m = np.ones(N) * (totmass / N)
    pos, vel = force_centers(pos, vel, center_pos, center_vel, force_origin)
    return pos, vel, m


def _hernquist_fE(E):
    E = np.asarray(E, dtype=np.float64)
    return (np.sqrt(E) * (1 - 2. * E) * (8. * E * E - 8 * E - 3.) / ((1. - E) ** 2) +
            3. * np.arcsin(np.sqrt(E)) / ((1. - E) ** (5. / 2))) / (8. * np.sqrt(2) * np.pi ** 3)


def Hernquist(N, a, totmass, cutoff=10., center_pos=None, center_vel=None, force_origin=True,
              seed=None):
    """Isotropic Hernquist sphere; a in kpc, totmass in Msun (gravhopper.py:1544-1605)."""
    from scipy import integrate
    rng = np.random.default_rng(seed)
    xi_cutoff = cutoff ** 2 / ((1. + cutoff) ** 2)
    rad_xi = rng.uniform(0.0, xi_cutoff, size=N)
    r_over_a = 1. / (1. / np.sqrt(rad_xi) - 1.)
    radius = r_over_a * a
    pos = _sphere(rng, radius, N)

    Eax = np.arange(0.0, 1.0, 0.002)
    cumulative_fE = [integrate.quad(_hernquist_fE, 0.0, Etop)[0] for Etop in Eax]
    potential = -1. / (1. + r_over_a)
    most_bound_potential = np.max(-potential)
    if most_bound_potential > Eax.max():
        # the reference discards this append (gravhopper.py:1573-1575) and then fails
        Eax = np.append(Eax, most_bound_potential)
        cumulative_fE.append(integrate.quad(_hernquist_fE, 0.0, most_bound_potential)[0])
    cumulative_fE = np.array(cumulative_fE) / np.max(cumulative_fE)
    Einterp = _interp_monotone(cumulative_fE, Eax)
    inverse_Einterp = _interp_monotone(Eax, cumulative_fE)

    max_possible_xi = inverse_Einterp(-potential)
    E_xi = rng.uniform(0.0, max_possible_xi, size=N)
    bindingE = Einterp(E_xi)
    energy_units = G * totmass / a
    velocity = np.sqrt(2. * np.maximum(-(bindingE + potential), 0.0) * energy_units)
    vel = _sphere(rng, velocity, N)
    m = np.ones(N) * (totmass / N)
    pos, vel = force_centers(pos, vel, center_pos, center_vel, force_origin)
    return pos, vel, m


def hernquist_vcirc(a, totmass):
    """Circular-velocity curve v_c(R [kpc]) [km/s] of a Hernquist halo (for expdisk)."""
    def vc(R):
        R = np.asarray(R, dtype=np.float64)
        return np.sqrt(G * totmass * R) / (R + a)
    return vc


def expdisk(N, sigma0, Rd, z0, sigmaR_Rd, external_rotcurve=None, center_pos=None,
            center_vel=None, force_origin=True, seed=None):
    """Exponential disk; sigma0 in Msun/kpc^2, Rd and z0 in kpc, sigmaR_Rd in km/s,
    external_rotcurve(R kpc) -> km/s (gravhopper.py:1669-1733)."""
    from scipy import special
    rng = np.random.default_rng(seed)
    totmass = np.pi * Rd ** 2 * sigma0
    Rax = np.arange(0.001 * Rd, 10 * Rd, 0.01 * Rd)
    R_cumprob = Rd ** 2 - Rd * np.exp(-Rax / Rd) * (Rax + Rd)
    R_cumprob /= R_cumprob[-1]
    # the reference's table starts at cumprob 5e-7 > 0 and raises for smaller deviates
    # (gravhopper.py:1675-1681); anchor it at (0, 0) instead.
    probtransform = _interp_monotone(np.concatenate(([0.0], R_cumprob)),
                                     np.concatenate(([0.0], Rax)))
    R_xi = rng.uniform(0.0, 1.0, size=N)
    R = np.maximum(probtransform(R_xi), 1e-6 * Rd)
    phi = rng.uniform(0.0, 2.0 * np.pi, size=N)
    x = R * np.cos(phi)
    y = R * np.sin(phi)
    z_xi = rng.uniform(0, 1.0, size=N)
    z = 2 * z0 * np.arctanh(z_xi)
    z *= (2 * (rng.uniform(0, 1, size=N) < 0.5)) - 1

    def om2(rad):
        y_R = rad / (2. * Rd)
        omega2 = np.pi * G * sigma0 / Rd * (special.iv(0, y_R) * special.kv(0, y_R) -
                                            special.iv(1, y_R) * special.kv(1, y_R))
        if external_rotcurve is not None:
            omega2 = omega2 + (external_rotcurve(rad) / rad) ** 2
        return omega2

    Omega2 = om2(R)
    h = 1e-3
    dom2 = (om2(R + h) - om2(np.maximum(R - h, 1e-9))) / (R + h - np.maximum(R - h, 1e-9))
    kappa2 = 4. * Omega2 + R * dom2
    sigma_R = sigmaR_Rd * np.exp(0.25 * (1 - R / Rd))
    sigma2_phi = sigma_R ** 2 * 4 * Omega2 / kappa2
    sigma2_z = np.pi * G * z0 * sigma0 * 0.5 * np.exp(-R / Rd)
    vphi_mean = R * np.sqrt(Omega2)
    vphi = np.sqrt(np.maximum(sigma2_phi, 0.0)) * rng.normal(size=N) + vphi_mean
    vR = sigma_R * rng.normal(size=N)
    vx = -vphi * np.sin(phi) + vR * np.cos(phi)
    vy = vphi * np.cos(phi) + vR * np.sin(phi)
    vz = np.sqrt(sigma2_z) * rng.normal(size=N)
    m = np.ones(N) * (totmass / N)
    pos, vel = force_centers(np.vstack((x, y, z)).T, np.vstack((vx, vy, vz)).T,
                             center_pos, center_vel, force_origin)
    return pos, vel, m


def TSIS(N, maxrad, totmass, center_pos=None, center_vel=None, force_origin=True, seed=None):
    """Truncated singular isothermal sphere (gravhopper.py:1378-1398)."""
    rng = np.random.default_rng(seed)
    sigma = np.sqrt(totmass * G / (2 * maxrad))
    radius = rng.uniform(0.0, maxrad, size=N)
    pos = _sphere(rng, radius, N)
    vx = rng.normal(0.0, sigma, size=N)
    vy = rng.normal(0.0, sigma, size=N)
    vz = rng.normal(0.0, sigma, size=N)
    m = np.ones(N) * (totmass / N)
    pos, vel = force_centers(pos, np.vstack((vx, vy, vz)).T, center_pos, center_vel,
                             force_origin)
    return pos, vel, m


def galaxy_model(N, disk_fraction=0.2, seed=1234):
    """BASELINE config 5 analogue: expdisk(sigma0=200 Msun/pc^2, Rd=2, z0=0.2 kpc,
    sigmaR(Rd)=20 km/s, in the halo's rotation curve) + Hernquist(a=20 kpc, 4e11 Msun)."""
    Nd = int(round(N * disk_fraction))
    Nh = N - Nd
    halo_a, halo_M = 20.0, 4e11
    pd, vd, md = expdisk(Nd, 200.0 * 1e6, 2.0, 0.2, 20.0,
                         external_rotcurve=hernquist_vcirc(halo_a, halo_M), seed=seed)
    ph, vh, mh = Hernquist(Nh, halo_a, halo_M, seed=seed + 1)
    return (np.vstack((pd, ph)), np.vstack((vd, vh)), np.hstack((md, mh)))
