"""Host-side logic that needs no GPU: the ty_gas_concs mirror, array layout helpers, the synthetic
generator's shard reproducibility, and the N>1 column sharding over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch

import ecckd_b200
from ecckd_b200 import sharding, synthetic
from ecckd_b200.gas_optics import PreparedGases, _fstrides, fortran_empty, to_fortran


def test_gas_concs_mirror():
    gc = ecckd_b200.GasConcs(["CO2 ", "h2o", "O3"])
    assert gc.get_num_gases() == 3 and gc.get_gas_names() == ["co2", "h2o", "o3"]   # lower-cased like ty_gas_concs%init
    gc.set_vmr("co2", 4e-4)
    gc.set_vmr("h2o", np.ones((4, 3)))
    with pytest.raises(ecckd_b200.EcckdError) as ei:
        gc.items()
    assert "gas o3 not found" in str(ei.value)
    gc.set_vmr("o3", np.arange(3.0))
    assert [n for n, _ in gc.items()] == ["co2", "h2o", "o3"]


def test_prepared_gases_views_follow_get_vmr_broadcasting():
    ncol, nlay = 4, 3
    pg = PreparedGases([("co2", 4e-4), ("o3", np.arange(3.0)), ("h2o", np.ones((ncol, nlay)))], ncol, nlay, np.float64)
    v = pg.views
    assert (v[0].col_stride, v[0].lay_stride) == (0, 0)          # scalar
    assert (v[1].col_stride, v[1].lay_stride) == (0, 1)          # (nlay) profile
    assert (v[2].col_stride, v[2].lay_stride) == (1, ncol)       # (ncol, nlay) field, column-major
    with pytest.raises(ecckd_b200.EcckdError):
        PreparedGases([("o3", np.arange(5.0))], ncol, nlay, np.float64)
    with pytest.raises(ecckd_b200.EcckdError):
        PreparedGases([("h2o", np.ones((ncol + 1, nlay)))], ncol, nlay, np.float64)


def test_fortran_layout_helpers():
    a = fortran_empty((5, 7, 3), np.float64)
    assert a.flags.f_contiguous and a.shape == (5, 7, 3)
    t = fortran_empty((5, 7, 3), torch.float64, "cpu")
    assert tuple(t.shape) == (5, 7, 3) and tuple(t.stride()) == _fstrides((5, 7, 3)) == (1, 5, 35)
    x = torch.arange(12.0, dtype=torch.float64).reshape(3, 4)
    y = to_fortran(x, torch.float64, "cpu")
    assert torch.equal(x, y) and tuple(y.stride()) == (1, 3)


def test_synthetic_atmosphere_is_physical_and_toa_first():
    atm = synthetic.make_atmosphere(64, 137)
    assert tuple(atm.plev.shape) == (64, 138) and tuple(atm.plev.stride()) == (1, 64)
    assert (atm.plev[:, 1:] > atm.plev[:, :-1]).all() and atm.plev[:, 0].eq(1.0).all()
    assert atm.plev[:, -1].min() >= 5.0e4 and atm.plev[:, -1].max() <= 1.05e5
    assert atm.tlay.min() >= 160 and atm.tlay.max() <= 320
    assert atm.h2o.min() >= 1e-7 and atm.h2o.max() <= 4e-2 and (atm.h2o == 1e-7).any()    # reaches below mole_fraction(1)
    assert [n for n, _ in atm.gas_list()] == list(synthetic.GAS_ORDER)


def test_synthetic_noise_scale_only_touches_the_noisy_fields():
    """`noise` scales the cell-by-cell white noise of SURVEY 8d config 4 (tlay, h2o); the default is the specification,
    0 removes it (bench.py's smooth_atmosphere extra) and leaves pressures, ozone and the scalars alone."""
    a, b = synthetic.make_atmosphere(12, 137), synthetic.make_atmosphere(12, 137, noise=0.0)
    assert torch.equal(a.plev, b.plev) and torch.equal(a.o3, b.o3) and a.scalars == b.scalars
    assert not torch.equal(a.tlay, b.tlay) and not torch.equal(a.h2o, b.h2o)
    # without noise a column's temperature is a smooth function of pressure: neighbouring layers differ by < 10 K
    # (with the N(0, 5 K) noise of the specification by more than 15 K somewhere in 12 x 137 cells)
    step = lambda t: float((t[:, 1:] - t[:, :-1]).abs().max())
    assert step(b.tlay) < 10.0 < 15.0 < step(a.tlay)
    assert torch.equal(synthetic.make_atmosphere(12, 137, noise=1.0).tlay, a.tlay)


def test_synthetic_columns_do_not_depend_on_the_shard():
    full = synthetic.make_atmosphere(100, 60)
    for world in (2, 3, 8):
        parts = [synthetic.make_atmosphere(c1 - c0, 60, col0=c0) for c0, c1 in
                 (sharding.shard_columns(100, world, r) for r in range(world)) if c1 > c0]
        assert torch.equal(torch.cat([p.plev for p in parts], 0), full.plev)
        # the random streams are functions of the global column index; torch's CPU log/cos/pow differ in the
        # last bit between SIMD body and scalar tail, hence "equal to rounding" for the derived fields
        for name in ("tlay", "tlev", "h2o", "o3"):
            torch.testing.assert_close(torch.cat([getattr(p, name) for p in parts], 0), getattr(full, name), rtol=1e-13, atol=0)
        torch.testing.assert_close(torch.cat([p.tsfc for p in parts]), full.tsfc, rtol=1e-13, atol=0)


def test_shard_columns_partitions_exactly():
    for ncol in (0, 1, 7, 100, 6_600_000):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_columns(ncol, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == ncol
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(c1 - c0 for c0, c1 in spans) - min(c1 - c0 for c0, c1 in spans) <= (ncol + world - 1) // world
    with pytest.raises(ValueError):
        sharding.shard_columns(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, ncol, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        from conftest import FILES  # noqa: F401  (sys.path side effect in spawned process)
        from oracle import ckd_model
        from oracle.c_oracle import COracle
        c0, c1 = sharding.shard_columns(ncol, world, rank)
        atm = synthetic.make_atmosphere(c1 - c0, 20, col0=c0)
        F = lambda t: np.asfortranarray(t.numpy())
        gl = [(n, F(v) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
        co = COracle(ckd_model.load_and_init(FILES["lw_fsck"]))
        out, err = co.gas_optics_int(F(atm.plev), F(atm.tlay), atm.tsfc.numpy(), gl, F(atm.tlev))
        # no data-path collective: each rank owns its columns; only timing / bookkeeping is reduced
        sharding.barrier()
        t_max = sharding.max_over_ranks(float(rank + 1))
        n_tot = sharding.sum_over_ranks(float(c1 - c0))
        gathered = [None] * world
        dist.all_gather_object(gathered, (c0, c1, float(out["tau"].sum()), out["tau"][:, 3, 5].copy()))
        if rank == 0:
            q.put((t_max, n_tot, gathered))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_column_sharding(ck_models, c_oracles):
    """N>1 path on CPU: two ranks each compute their column shard; the union equals the single-rank result."""
    import torch.multiprocessing as mp
    ncol, world = 21, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, ncol, q)) for r in range(world)]
    for p in procs:
        p.start()
    t_max, n_tot, gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == 2.0 and n_tot == ncol
    atm = synthetic.make_atmosphere(ncol, 20)
    F = lambda t: np.asfortranarray(t.numpy())
    gl = [(n, F(v) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
    out, _ = c_oracles["lw_fsck"].gas_optics_int(F(atm.plev), F(atm.tlay), atm.tsfc.numpy(), gl, F(atm.tlev))
    cols = np.concatenate([g[3] for g in sorted(gathered)])
    from conftest import rel_err
    assert rel_err(cols, out["tau"][:, 3, 5]) <= 1e-12    # sharded == unsharded (inputs equal to rounding, see above)
    assert [g[:2] for g in sorted(gathered)] == [(0, 11), (11, 21)]


def _walk_plan(ncol, shortwave, sms):
    import ctypes as C
    lib = ecckd_b200.load_library()
    tpc, blocks = C.c_int(0), C.c_int(0)
    lib.ecckd_debug_walk_plan(ncol, int(shortwave), sms, C.byref(tpc), C.byref(blocks))
    return tpc.value, blocks.value


@pytest.mark.parametrize("sms", [1, 4, 148])
@pytest.mark.parametrize("shortwave", [False, True])
def test_walk_plan_covers_every_tile_exactly_once(shortwave, sms):
    """The persistent walk kernel's launch geometry (host arithmetic, no device): block b of `blocks` walks chunks
    b, b + blocks, ... of `tiles_per_chunk` 16-column tiles.  Replay the kernel's tile_of() enumeration and check the
    partition; also the policy: two-tile chunks only in the longwave, only for large calls, never unbalancing."""
    for ncol in [1, 15, 16, 17, 70, 1800, 2368, 9472, 9473, 16384, 18921, 18922, 32768, 65536, 65535, 100000, 1048576]:
        tpc, blocks = _walk_plan(ncol, shortwave, sms)
        ntile = (ncol + 15) // 16
        nchunk_all = (ntile + tpc - 1) // tpc
        assert tpc in (1, 2) and 1 <= blocks <= max(1, min(sms, nchunk_all))
        if shortwave or ntile < 4 * sms:
            assert tpc == 1
        seen = np.zeros(ntile, dtype=np.int32)
        longest = 0
        for b in range(blocks):
            nchunk = (nchunk_all - b + blocks - 1) // blocks          # as in walk_kernel
            assert nchunk >= 1
            mine = 0
            for ts in range(nchunk * tpc):
                k = ts // tpc
                gt = (b + k * blocks) * tpc + (ts - k * tpc)
                if gt < ntile:
                    seen[gt] += 1
                    mine += 1
            longest = max(longest, mine)
        assert (seen == 1).all(), f"ncol={ncol}: tiles walked {seen.min()}..{seen.max()} times"
        # balance: the longest block is within one chunk of the ideal share
        assert longest <= -(-ntile // blocks) + tpc - 1 + (tpc - 1)


def test_constant_divisor_division_in_numpy(ck_models):
    """The device divides by per-model constants as q0 = x*RN(1/c), q = fma(fma(-c, q0, x), RN(1/c), q0) (div_const in
    csrc/ecckd_device.cuh).  The same sequence, with the FMAs evaluated exactly in Python fractions and rounded once,
    must return the IEEE quotient x / c for every divisor of the three ck files (the GPU test repeats this on the device
    with 2^22 dividends per divisor)."""
    from fractions import Fraction
    rng = np.random.default_rng(11)

    def fma(a, b, c):        # one rounding: exact in rationals, then to the nearest double
        return float(Fraction(a) * Fraction(b) + Fraction(c))
    divisors = set()
    for m in ck_models.values():
        divisors.add(float(m.log_pressure[1] - m.log_pressure[0]))
        divisors.add(float(m.temperature[1, 0] - m.temperature[0, 0]))
        if not m.shortwave:
            divisors.add(float(m.temperature_planck[1] - m.temperature_planck[0])); divisors.add(float(m.temperature_planck[0]))
        for tab in m.absorption:
            if tab.mole_fraction is not None:
                divisors.add(float(np.log(tab.mole_fraction[1] / tab.mole_fraction[0])))
    divisors.add(float(np.float64(np.float32(3.14159265359))))
    for c in sorted(divisors):
        rc = 1.0 / c
        xs = np.concatenate([(rng.random(400) - 0.5) * 60.0, (rng.random(400) - 0.3) * 400.0,
                             c * rng.integers(1, 64, 400) * (1.0 + rng.integers(-3, 4, 400) * 2.220446049250313e-16)])
        for x in xs:
            x = float(x)
            q0 = x * rc
            q = fma(fma(-c, q0, x), rc, q0)
            assert q == x / c, (c, x)
