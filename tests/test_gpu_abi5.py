"""ABI 5 features of the CUDA path: broadcast rows (and the folded-geometry kernels they select),
compact and float32 results, the staged host path for pageable memory, input validation, bare-soil
(LAI = 0) hot-spot integrals, user thermal leaf optics, context caching."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import load_golden, relerr  # noqa: E402

RTOL64 = 1e-9
RTOL32 = 1e-4


@pytest.fixture(scope="module")
def env():
    import torch
    import spart_b200
    import spart_oracle as so
    assert torch.cuda.is_available()
    return torch, spart_b200, so


def _dev(torch, P, dtype=np.float64):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(P, dtype=np.float64).T.astype(dtype))).cuda()


def test_bare_soil_lai_zero(env):
    """LAI = 0 with a narrow hot spot (the analytic remainder of the hot-spot integral divided
    0 by 0): the reference returns finite bare-soil values (sailh.py:112-114 guards LAI > 0)."""
    torch, sb, so = env
    P = so.synthetic_params(32, 3, seed=5)
    P[:16, so.LAI] = 0.0
    P[16:, so.LAI] = [1e-12, 1e-9, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 0.05] * 2
    P[:, so.HOT_Q] = np.tile([0.01, 0.05, 0.001, 0.2], 8)
    P[:8, so.SZA], P[:8, so.VZA], P[:8, so.RAA] = 40.0, 0.0, 0.0
    for sensor in ("LANDSAT8-OLI", "Sentinel2A-MSI"):
        want = so.spart_bands(P, sensor)
        assert np.isfinite(want).all()
        got = sb.run_batch_params(_dev(torch, P), sensor).cpu().numpy()
        assert np.isfinite(got).all()
        assert relerr(got, want) < RTOL64
        got32 = sb.run_batch_params(_dev(torch, P), sensor, precision="fp32").cpu().numpy()
        assert np.isfinite(got32).all()
        assert relerr(got32, want) < RTOL32
    g = load_golden("edge_lai0.npz")          # the unmodified reference on bare-soil rows
    got = sb.run_batch_params(_dev(torch, g["params"]), str(g["sensor"])).cpu().numpy()
    assert relerr(got, g["O2"]) < RTOL64 and relerr(got, g["O1"]) < 5e-8


def test_broadcast_rows_equal_full_rows(env):
    """A broadcast row is read once: the result must be the bits of the same batch with the value
    replicated (general kernels), and garbage beyond element 0 of such a row must not matter."""
    torch, sb, so = env
    P = so.synthetic_params(3000, 3, seed=31)
    P[:, so.SMC], P[:, so.FILM], P[:, so.DOY], P[:, so.NSTRUCT] = 25.0, 0.015, 123.0, 1.7
    P[:, so.LIDFA], P[:, so.LIDFB] = -0.3, 0.1
    rows = [so.SMC, so.FILM, so.DOY, so.NSTRUCT, so.LIDFA, so.LIDFB]
    full = sb.run_batch_params(_dev(torch, P), "LANDSAT8-OLI").cpu().numpy()
    Q = P.copy()
    Q[1:, rows] = np.nan
    dev = _dev(torch, Q)
    got = sb.run_batch_params(dev, "LANDSAT8-OLI", broadcast_rows=rows)
    assert np.array_equal(got.cpu().numpy(), full)
    host = sb.run_batch_params(np.ascontiguousarray(Q.T), "LANDSAT8-OLI", broadcast_rows=rows)
    assert np.array_equal(host, full)
    got32 = sb.run_batch_params(dev, "LANDSAT8-OLI", broadcast_rows=rows, precision="fp32").cpu().numpy()
    assert np.array_equal(got32, sb.run_batch_params(_dev(torch, P), "LANDSAT8-OLI", precision="fp32").cpu().numpy())
    with pytest.raises(ValueError):
        sb.run_batch_params(dev, "LANDSAT8-OLI", broadcast_rows=[27])
    # LIDFa and LIDFb both broadcast: the leaf-angle iterations run once for the batch (also with a shared
    # geometry); only one of them broadcast: the general path
    P[:, so.SZA], P[:, so.VZA], P[:, so.RAA] = 33.0, 7.0, 40.0
    full = sb.run_batch_params(_dev(torch, P), "Sentinel2A-MSI").cpu().numpy()
    for rows in ([so.LIDFA, so.LIDFB, so.SZA, so.VZA, so.RAA], [so.LIDFA, so.SZA, so.VZA, so.RAA], [so.LIDFB]):
        Q = P.copy()
        Q[1:, rows] = np.nan
        got = sb.run_batch_params(_dev(torch, Q), "Sentinel2A-MSI", broadcast_rows=rows).cpu().numpy()
        if so.SZA in rows:      # the folded-geometry kernels re-associate a few sums
            assert relerr(got, full) < 1e-12
        else:
            assert np.array_equal(got, full)
    a = sb.run_batch_params(_dev(torch, P), "Sentinel2A-MSI", broadcast_rows=[so.SZA, so.VZA, so.RAA]).cpu().numpy()
    Q = P.copy()
    Q[1:, [so.LIDFA, so.LIDFB]] = np.nan
    b = sb.run_batch_params(_dev(torch, Q), "Sentinel2A-MSI",
                            broadcast_rows=[so.LIDFA, so.LIDFB, so.SZA, so.VZA, so.RAA]).cpu().numpy()
    assert np.array_equal(a, b)


def test_uniform_geometry_is_validated(env):
    """uniform_geometry=True is checked (it was an unchecked promise): varying angles raise,
    on the device path and on the host path; angle rows passed as broadcast rows cannot vary."""
    torch, sb, so = env
    P = so.synthetic_params(2000, 2, seed=8)
    dev = _dev(torch, P)
    a = sb.run_batch_params(dev, "Sentinel2A-MSI", uniform_geometry=True).cpu().numpy()
    b = sb.run_batch_params(dev, "Sentinel2A-MSI", broadcast_rows=[19, 20, 21]).cpu().numpy()
    assert np.array_equal(a, b)
    P[1234, so.VZA] = 1.0
    with pytest.raises(sb.SpartError):
        sb.run_batch_params(_dev(torch, P), "Sentinel2A-MSI", uniform_geometry=True)
    with pytest.raises(sb.SpartError):
        sb.run_batch_params(np.ascontiguousarray(P.T), "Sentinel2A-MSI", uniform_geometry=True)
    with pytest.raises(sb.SpartError):
        sb.run_batch_params(_dev(torch, P), "Sentinel2A-MSI", uniform_geometry=True, precision="fp32")
    # broadcast rows: element 0 defines the geometry, whatever else the row holds
    c = sb.run_batch_params(_dev(torch, P), "Sentinel2A-MSI", broadcast_rows=[19, 20, 21]).cpu().numpy()
    assert np.array_equal(c, a)


@pytest.mark.parametrize("sensor,cfg,uniform", [("Sentinel2A-MSI", 2, True), ("LANDSAT8-OLI", 3, False),
                                                ("TerraAqua-MODIS", 3, False), ("Sentinel3A-OLCI", 2, True)])
def test_compact_output_rebuilds_l_toa_bit_exactly(env, sensor, cfg, uniform):
    torch, sb, so = env
    P = so.synthetic_params(3333, cfg, seed=12)
    dev = _dev(torch, P)
    full = sb.run_batch_params(dev, sensor, uniform_geometry=uniform)
    c = sb.run_batch_params(dev, sensor, uniform_geometry=uniform, compact=True)
    assert isinstance(c, sb.CompactBands) and c.R.shape == (3333, full.shape[1], 2) and c.etscale.shape == (3333,)
    assert torch.equal(c.R, full[..., :2])
    assert torch.equal(c.L_TOA, full[..., 2])
    assert torch.equal(c.full(), full)
    h = sb.run_batch_params(np.ascontiguousarray(P.T), sensor, uniform_geometry=uniform, compact=True)
    assert np.array_equal(h.full(), full.cpu().numpy())
    # SRF band mode and FP32 mode write the same layout
    s = sb.run_batch_params(dev[:, :500].contiguous(), sensor, band_mode="srf", compact=True)
    assert torch.equal(s.full(), sb.run_batch_params(dev[:, :500].contiguous(), sensor, band_mode="srf"))
    f = sb.run_batch_params(dev, sensor, precision="fp32", compact=True)
    assert torch.equal(f.full(), sb.run_batch_params(dev, sensor, precision="fp32"))


@pytest.mark.parametrize("sensor,cfg", [("Sentinel2A-MSI", 2), ("LANDSAT8-OLI", 3)])
def test_f32_io(env, sensor, cfg):
    """SPART_FLAG_F32_IO: float32 parameters in, float32 results out.  The model is evaluated at the
    float-rounded inputs, so the oracle gets those; the gate is the FP32-mode tolerance."""
    torch, sb, so = env
    P = so.synthetic_params(20000, cfg, seed=40 + cfg)
    if cfg == 3:
        P[:, so.RAA] = np.round(P[:, so.RAA])      # cos(rel * 180/pi) (smac.py:130) amplifies float rounding of rel
    P32 = P.astype(np.float32)
    want, canopy = so.spart_bands(P32.astype(np.float64), sensor, return_canopy=True)
    valid = ((canopy > 0) & (canopy < 1)).all(axis=2)
    dev = torch.from_numpy(np.ascontiguousarray(P32.T)).cuda()
    got = sb.run_batch_params(dev, sensor, precision="fp32")
    assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
    e = np.abs(got.cpu().numpy().astype(np.float64) - want) / np.abs(want)
    assert e[valid].max() < 1.5e-4                 # FP32 arithmetic (1e-4) + float32 storage of the result (6e-8)
    # same arithmetic as FP32 mode on double-typed copies of the same numbers
    d = sb.run_batch_params(dev.double(), sensor, precision="fp32")
    assert torch.equal(got, d.float())
    host = sb.run_batch_params(np.ascontiguousarray(P32.T), sensor, precision="fp32")
    assert host.dtype == np.float32 and np.array_equal(host, got.cpu().numpy())
    c = sb.run_batch_params(dev, sensor, precision="fp32", compact=True)
    assert c.buf.dtype == torch.float32 and torch.equal(c.full(), got)
    with pytest.raises(ValueError):
        sb.run_batch_params(dev, sensor)           # float32 parameters need precision="fp32"


def test_host_path_pageable_and_pinned(env):
    """Pageable NumPy arrays are staged by the library's copy threads, pinned tensors are DMA'd
    directly; both must give the device path's bits, for several chunks and a ragged tail."""
    torch, sb, so = env
    n = 3 * 65536 + 777
    P = so.synthetic_params(n, 2, seed=3)
    pt = np.ascontiguousarray(P.T)
    dev = sb.run_batch_params(torch.from_numpy(pt).cuda(), "Sentinel2A-MSI", uniform_geometry=True).cpu().numpy()
    pageable = sb.run_batch_params(pt, "Sentinel2A-MSI", uniform_geometry=True)
    assert np.array_equal(pageable, dev)
    pin_in = torch.from_numpy(pt).pin_memory()
    pin_out = torch.empty((n, 13, 3), dtype=torch.float64).pin_memory()
    sb.run_batch_params(pin_in, "Sentinel2A-MSI", out=pin_out, uniform_geometry=True)
    assert np.array_equal(pin_out.numpy(), dev)
    mixed = sb.run_batch_params(pin_in, "Sentinel2A-MSI", uniform_geometry=True)          # pinned in, pageable out
    assert np.array_equal(mixed, dev)
    out2 = torch.empty((n, 13, 3), dtype=torch.float64).pin_memory()
    sb.run_batch_params(pt, "Sentinel2A-MSI", out=out2, broadcast_rows=[7, 8, 13, 14, 19, 20, 21])
    assert np.array_equal(out2.numpy(), dev)
    c = sb.run_batch_params(pt, "Sentinel2A-MSI", uniform_geometry=True, compact=True)
    assert np.array_equal(c.full(), dev)


def test_host_path_many_spans_and_chunks(env, monkeypatch):
    """Small spans / chunks (SPART_HOST_SPAN / SPART_HOST_CHUNK) so that the three span buffers and the four
    output slots are reused several times within one call: pinned and pageable buffers, full / compact / float32
    results, a ragged tail, broadcast rows (their elements are written once per call and buffer)."""
    torch, sb, so = env
    monkeypatch.setenv("SPART_HOST_SPAN", "4096")
    monkeypatch.setenv("SPART_HOST_CHUNK", "1024")
    n = 11 * 4096 + 1234 + 77
    P = so.synthetic_params(n, 2, seed=21)
    P[:, so.LIDFA], P[:, so.LIDFB] = -0.35, -0.15
    rows = [7, 8, 13, 14, so.LIDFA, so.LIDFB, 19, 20, 21]
    pt = np.ascontiguousarray(P.T)
    Q = pt.copy()
    Q[rows, 1:] = np.nan                                   # only element 0 of a broadcast row may be read
    dev = sb.run_batch_params(torch.from_numpy(pt).cuda(), "Sentinel2A-MSI", broadcast_rows=rows).cpu().numpy()
    assert np.array_equal(sb.run_batch_params(Q, "Sentinel2A-MSI", broadcast_rows=rows), dev)
    pin_in = torch.from_numpy(Q).pin_memory()
    pin_out = torch.empty((n, 13, 3), dtype=torch.float64).pin_memory()
    sb.run_batch_params(pin_in, "Sentinel2A-MSI", out=pin_out, broadcast_rows=rows)
    assert np.array_equal(pin_out.numpy(), dev)
    c = sb.run_batch_params(pin_in, "Sentinel2A-MSI", broadcast_rows=rows, compact=True)     # etscale once per span
    assert np.array_equal(np.asarray(c.full()), dev)
    c2 = sb.run_batch_params(Q, "Sentinel2A-MSI", broadcast_rows=rows, compact=True)
    assert np.array_equal(np.asarray(c2.full()), dev)
    dev32 = sb.run_batch_params(torch.from_numpy(pt.astype(np.float32)).cuda(), "Sentinel2A-MSI", precision="fp32",
                                broadcast_rows=rows).cpu().numpy()
    h32 = sb.run_batch_params(torch.from_numpy(Q.astype(np.float32)).pin_memory(), "Sentinel2A-MSI",
                              precision="fp32", broadcast_rows=rows, compact=True)
    assert np.array_equal(np.asarray(h32.full()), dev32)
    # no broadcast rows at all, general geometry, a second sensor
    P3 = so.synthetic_params(9000, 3, seed=22)
    p3 = np.ascontiguousarray(P3.T)
    d3 = sb.run_batch_params(torch.from_numpy(p3).cuda(), "LANDSAT8-OLI").cpu().numpy()
    assert np.array_equal(sb.run_batch_params(p3, "LANDSAT8-OLI"), d3)


def test_thermal_leaf_optics_are_passed_through(env):
    """LeafBiology.rho_thermal / tau_thermal (SPART.py:461-466) reach leafopt and canopyopt."""
    torch, sb, so = env
    g = load_golden("spectra.npz")
    p = g["params"][0]
    mk = lambda leaf: sb.SPART(sb.SoilParameters(*p[9:15]), leaf, sb.CanopyStructure(*p[15:19]),
                               sb.AtmosphericProperties(*p[22:26]), sb.Angles(*p[19:22]), "Sentinel2A-MSI", int(p[26]))
    a = mk(sb.LeafBiology(*p[0:9]))
    b = mk(sb.LeafBiology(*p[0:9], rho_thermal=0.03, tau_thermal=0.02))
    a.run(), b.run()
    assert (a.leafopt.refl[2001:, 0] == 0.01).all() and (b.leafopt.refl[2001:, 0] == 0.03).all()
    assert (b.leafopt.tran[2001:, 0] == 0.02).all()
    assert np.array_equal(a.canopyopt.rso[:2001], b.canopyopt.rso[:2001])
    assert not np.allclose(a.canopyopt.rso[2001:], b.canopyopt.rso[2001:])


def test_user_assigned_lidf(env):
    """An assigned leaf inclination distribution replaces the one derived from LIDFa / LIDFb, as in the
    reference (canopy.lidf = ...; sailh.py:81-97 read canopy.lidf): golden rows recorded from the unmodified
    reference, a batch against the oracle, the SPART class, and the spectra."""
    torch, sb, so = env
    g = load_golden("user_lidf.npz")
    P, L, sensor = g["params"], g["lidf"], str(g["sensor"])
    got = sb.run_batch_params(_dev(torch, P), sensor, lidf=torch.from_numpy(L).cuda()).cpu().numpy()
    assert relerr(got, g["O2"]) < RTOL64 and relerr(got, g["O1"]) < 5e-8
    host = sb.run_batch_params(np.ascontiguousarray(P.T), sensor, lidf=L)                 # NumPy in, NumPy out
    assert np.array_equal(host, got)
    # a batch: the distribution the kernels would derive themselves, passed explicitly, and a random one
    P2 = so.synthetic_params(3000, 2, seed=77)
    own = so.leafangles(P2[:, so.LIDFA], P2[:, so.LIDFB])
    a = sb.run_batch_params(_dev(torch, P2), "Sentinel2A-MSI", uniform_geometry=True).cpu().numpy()
    b = sb.run_batch_params(_dev(torch, P2), "Sentinel2A-MSI", uniform_geometry=True, lidf=own).cpu().numpy()
    assert relerr(b, a) < 1e-9
    rnd = np.random.default_rng(5).dirichlet(np.ones(13), size=3000)
    c = sb.run_batch_params(_dev(torch, P2), "Sentinel2A-MSI", uniform_geometry=True, lidf=rnd, compact=True)
    assert relerr(c.full().cpu().numpy(), so.spart_bands(P2, "Sentinel2A-MSI", lidf=rnd)) < RTOL64
    one = sb.run_batch_params(_dev(torch, P2), "Sentinel2A-MSI", lidf=rnd[7]).cpu().numpy()  # one row for the batch
    assert relerr(one, so.spart_bands(P2, "Sentinel2A-MSI", lidf=np.tile(rnd[7], (3000, 1)))) < RTOL64
    with pytest.raises(ValueError):
        sb.run_batch_params(_dev(torch, P2), "Sentinel2A-MSI", lidf=rnd, precision="fp32")
    # the class API
    p = P[3]
    canopy = sb.CanopyStructure(*p[15:19])
    canopy.lidf = L[3]
    assert canopy.lidf.shape == (13, 1) and np.array_equal(canopy.lidf[:, 0], L[3])
    sp = sb.SPART(sb.SoilParameters(*p[9:15]), sb.LeafBiology(*p[0:9]), canopy, sb.AtmosphericProperties(*p[22:26]),
                  sb.Angles(*p[19:22]), sensor, int(p[26]))
    df = sp.run()
    assert relerr(np.stack([df["R_TOC"], df["R_TOA"], df["L_TOA"]], axis=1), g["O2"][3]) < RTOL64
    plain = sb.SPART(sb.SoilParameters(*p[9:15]), sb.LeafBiology(*p[0:9]), sb.CanopyStructure(*p[15:19]),
                     sb.AtmosphericProperties(*p[22:26]), sb.Angles(*p[19:22]), sensor, int(p[26]))
    plain.run()
    assert not np.allclose(sp.canopyopt.rso[:2001], plain.canopyopt.rso[:2001])          # the spectra follow it too
    assert np.array_equal(sp.leafopt.refl, plain.leafopt.refl)
    with pytest.raises(ValueError):
        canopy.lidf = np.ones(12)


def test_sensor_contexts_keyed_by_content_and_bounded(env):
    torch, sb, so = env
    from spart_b200 import engine as E
    eng = sb.Engine()
    info = sb.load_sensor_info("LANDSAT8-OLI")
    P = so.synthetic_params(50, 3, seed=2)
    dev = _dev(torch, P)
    a = eng.forward_bands(dev, info).cpu().numpy()
    assert np.array_equal(a, eng.forward_bands(dev, "LANDSAT8-OLI").cpu().numpy())
    info["SMAC_coef"]["taur"] = info["SMAC_coef"]["taur"] * 1.5          # in-place edit of the same dict
    b = eng.forward_bands(dev, info).cpu().numpy()
    assert not np.array_equal(a, b)
    # the single-run class notices an edited shipped table as well
    s = sb.SPART(sb.SoilParameters(0.5, 0, 100, 20, 25, 0.015), sb.LeafBiology(40, 0.01, 0.02, 0, 10, 10, 1.5),
                 sb.CanopyStructure(3, -0.35, -0.15, 0.05), sb.AtmosphericProperties(0.325, 0.35, 1.41),
                 sb.Angles(40, 0, 0), "LANDSAT8-OLI", 100)
    r0 = s.run()["R_TOA"].to_numpy().copy()
    s.sensorinfo["SMAC_coef"]["taur"] = s.sensorinfo["SMAC_coef"]["taur"] * 1.5
    assert not np.allclose(s.run()["R_TOA"].to_numpy(), r0)
    # the cache is bounded: many distinct soil spectra do not accumulate contexts
    rng = np.random.default_rng(0)
    for _ in range(E.MAX_CONTEXTS + 5):
        eng.forward_bands(dev, "LANDSAT8-OLI", soil_spectrum=rng.uniform(0.1, 0.4, 2001))
    assert len(eng._ctx) <= E.MAX_CONTEXTS
    eng.close()


def test_two_gpus_in_one_process(env):
    """entry points must work whatever device is current and must not change it."""
    torch, sb, so = env
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    P = so.synthetic_params(500, 3, seed=6)
    want = so.spart_bands(P, "LANDSAT8-OLI")
    torch.cuda.set_device(0)
    d1 = _dev(torch, P).to("cuda:1")
    got1 = sb.run_batch_params(d1, "LANDSAT8-OLI")
    assert torch.cuda.current_device() == 0 and got1.device.index == 1
    assert relerr(got1.cpu().numpy(), want) < RTOL64
    got0 = sb.run_batch_params(_dev(torch, P), "LANDSAT8-OLI")
    assert torch.equal(got0.cpu(), got1.cpu())
    host1 = sb.run_batch_params(np.ascontiguousarray(P.T), "LANDSAT8-OLI", device=1)
    assert torch.cuda.current_device() == 0 and np.array_equal(host1, got0.cpu().numpy())
    spec1 = sb.default_engine(1).forward_spectrum(d1[:, :4].contiguous())
    assert torch.equal(spec1.cpu(), sb.default_engine(0).forward_spectrum(d1[:, :4].contiguous().to("cuda:0")).cpu())


def test_sharded_lut_single_rank_pieces(env):
    """The pieces of the sharded retrieval on one GPU: two slices searched with index offsets, their
    packed words min-reduced, equal the search over the whole table."""
    torch, sb, so = env
    from spart_b200 import lut
    g = torch.Generator(device="cuda").manual_seed(4)
    L = torch.rand((50_000, 13), generator=g, device="cuda", dtype=torch.float32)
    O = torch.rand((999, 13), generator=g, device="cuda", dtype=torch.float32)
    idx, cost = lut.nearest(L, O)
    w0 = lut._search(L[:20_000], O, None, 0, packed=True)
    w1 = lut._search(L[20_000:], O, None, 20_000, packed=True)
    i2, c2 = lut.unpack(torch.minimum(w0, w1))
    assert torch.equal(i2, idx) and torch.equal(c2, cost)


@pytest.mark.parametrize("nb,n,m", [(13, 50_000, 3000), (26, 20_000, 1000), (6, 4000, 777), (1, 300, 33), (20, 9999, 257)])
def test_lut_tensor_core_search(env, nb, n, m):
    """method="tensor" (3xTF32 mma.sync search + exact re-costing): the chosen entry is optimal up to the
    3xTF32 comparison error, the reported cost is the exact FP32 cost of that entry, self-matches are found."""
    torch, sb, so = env
    from spart_b200 import lut
    g = torch.Generator(device="cuda").manual_seed(nb)
    L = torch.rand((n, nb), generator=g, device="cuda", dtype=torch.float32)
    O = torch.rand((m, nb), generator=g, device="cuda", dtype=torch.float32)
    w = torch.rand(nb, generator=g, device="cuda", dtype=torch.float32) + 0.1
    for weights in (None, w):
        idx, cost = lut.nearest(L, O, weights, method="tensor")
        ie, ce = lut.nearest(L, O, weights)
        ww = torch.ones(nb, device="cuda") if weights is None else weights
        d = ((O[:, None, :].double() - L[None, :, :].double()) ** 2 * ww.double()).sum(-1)
        chosen = d.gather(1, idx[:, None])[:, 0]
        scale = ((O.double() ** 2 * ww.double()).sum(-1).sqrt()[:, None] * (L.double() ** 2 * ww.double()).sum(-1).sqrt()[None, :]).max()
        assert torch.all(chosen <= d.min(dim=1).values + 4e-6 * scale)
        assert torch.allclose(cost.double(), chosen, rtol=1e-4, atol=1e-9)          # the cost is that entry's exact cost
        assert (idx == ie).float().mean() > (0.999 if m >= 1000 else 0.9)     # near-ties may resolve differently
    idx, cost = lut.nearest(L, L[100:164].clone(), method="tensor")
    if nb >= 6:
        assert torch.equal(idx.cpu(), torch.arange(100, 164)) and float(cost.max()) == 0.0
    else:
        # few bands: other entries lie within the 3xTF32 resolution of a self-match (documented near-tie
        # behaviour of method="tensor"); the entry found is as good as the self-match up to that resolution
        assert float(cost.max()) <= 4e-6 * float((L ** 2).sum(-1).max())
        assert (idx.cpu() == torch.arange(100, 164)).float().mean() > 0.8
    w0 = lut._search(L[:n // 3], O, None, 0, packed=True, method="tensor")
    w1 = lut._search(L[n // 3:], O, None, n // 3, packed=True, method="tensor")
    i2, c2 = lut.unpack(torch.minimum(w0, w1))
    i1, c1 = lut.nearest(L, O, method="tensor")
    assert torch.equal(c2, c1)
