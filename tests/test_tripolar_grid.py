"""Host-side mirror of the reference's TripolarGrid (BASELINE config 5; /root/reference/src/OceanConfigurations/
one_degree_tripolar.jl:20-73): the geometry the flux path needs — fold adjacency, rotation angles, fractional indices — and the
oracle on it.  CPU only."""
import numpy as np

import climaocean.jl_b200 as cj
from tests.common import oracle_update


def _xyz(lam, phi):
    lam, phi = np.deg2rad(lam), np.deg2rad(phi)
    return np.stack([np.cos(phi) * np.cos(lam), np.cos(phi) * np.sin(lam), np.sin(phi)])


def test_fold_pairs_neighbouring_cells_and_rows_join_the_regular_part():
    g = cj.TripolarGrid((360, 180, 4))
    jj, ii = np.meshgrid(np.arange(g.Ny), np.arange(g.Nx), indexing="ij")
    lam, phi = g.coordinates(ii, jj)
    assert phi.max() < 90.0 and phi[:g.j0].max() < g.phi0 < phi[g.j0:].min() + 1.0
    # regular part: plain latitude–longitude rows
    assert np.allclose(lam[0], (np.arange(360) + 0.5 + g.lam_pole) % 360.0)
    # the top row folds onto itself: cell i and cell Nx-1-i are neighbours (closer than one meridional spacing)
    a = _xyz(lam[-1], phi[-1])
    d = np.degrees(np.arccos(np.clip((a * a[:, ::-1]).sum(0), -1, 1)))
    assert d.max() < 170.0 / 180.0
    # halo coordinates are those of the folded / periodic partner
    l1, p1 = g.coordinates(np.array([5]), np.array([g.Ny + 1]))
    l2, p2 = g.coordinates(np.array([g.Nx - 1 - 5]), np.array([g.Ny - 2]))
    assert l1 == l2 and p1 == p2
    l3, _ = g.coordinates(np.array([-2]), np.array([10]))
    l4, _ = g.coordinates(np.array([g.Nx - 2]), np.array([10]))
    assert l3 == l4


def test_rotation_is_a_rotation_and_trivial_below_the_cap():
    g = cj.TripolarGrid((360, 180, 4))
    cs, sn = g.rotation(ring=1)
    assert np.abs(cs ** 2 + sn ** 2 - 1).max() < 1e-14
    assert np.all(cs[0, 1:g.j0 - 1] > 1 - 1e-12)                       # regular rows: east is +i
    # on the fold the i-direction runs along the line between the two poles, i.e. due north on one half, due south on the other
    assert np.median(cs[0, -2, 1:-1]) < 0.1 and sn[0, -2, 1:-1].max() > 0.99 and sn[0, -2, 1:-1].min() < -0.99
    # north halo rows carry the frame of their folded partner, reversed
    assert np.allclose(cs[0, -1, 1:-1], -cs[0, -2, 1:-1][::-1]) and np.allclose(sn[0, -1, 1:-1], -sn[0, -2, 1:-1][::-1])


def test_oracle_on_the_tripolar_grid_rotates_the_winds():
    g = cj.TripolarGrid((72, 40, 2), halo=(4, 4, 2))
    host = cj.SurfaceFluxData.synthetic(g, ring=1)
    cfg = cj.default_config(72, 40, 2, 64)
    cfg.grid.ring = 1
    rot = dict(oracle_update(host, cfg))
    keep, host.rotation = host.rotation, None
    plain = dict(oracle_update(host, cfg))
    host.rotation = keep
    assert np.allclose(np.hypot(rot["exchange.u"], rot["exchange.v"]), np.hypot(plain["exchange.u"], plain["exchange.v"]), rtol=1e-12)
    assert np.array_equal(rot["exchange.u"][:g.j0 - 1], plain["exchange.u"][:g.j0 - 1])
    assert np.abs(rot["exchange.u"][g.j0 + 1:] - plain["exchange.u"][g.j0 + 1:]).max() > 1.0
    assert np.array_equal(rot["exchange.T"], plain["exchange.T"])        # scalars are not rotated
    # north halo of a scalar is the folded partner; vector components change sign
    T, u = host.ocean["T"].data, host.ocean["u"].data
    H, Nx, Ny = 4, 72, 40
    assert np.array_equal(T[:, H + Ny, H:H + Nx], T[:, H + Ny - 1, H:H + Nx][:, ::-1])
    assert np.array_equal(u[:, H + Ny + 1, H:H + Nx], -u[:, H + Ny - 2, H:H + Nx][:, ::-1])
