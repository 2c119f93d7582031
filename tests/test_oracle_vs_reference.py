"""CPU tests that run only where the reference itself has been compiled (oracle/_ref, built by
oracle/Makefile from /root/reference): the NumPy oracle against the live reference on cases the
golden files do not hold."""
import numpy as np
import pytest

import ref_bindings as rb
from dqmc_oracle import SdwOracle, SdwParams, HubbardOracle, HubbardParams
from dsfmt_oracle import RngOracle
from helpers import maxabs

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (make -C oracle)")


def test_rng_live():
    for seed, idx in ((99, 3), (1020304050, 17)):
        ref = rb.RefRng(seed, idx).draw(1200)
        r = RngOracle(seed, idx)
        assert np.array_equal(np.array([r.rand01() for _ in range(1200)]), ref)


@pytest.mark.parametrize("kw", [dict(L=4, m=12, s=5, rngIndex=9, r=0.4, mu=0.2),
                                dict(L=4, m=16, s=4, rngIndex=4, weakZflux=False, bc=2, delaySteps=1)])
def test_sdw_live(kw):
    p = SdwParams(**kw)
    o, r = SdwOracle(p), rb.RefSdw(p)
    assert maxabs(o.green[0], r.green()) < 1e-12
    for sw in range(4):
        o.sweep_thermalization()
        r.sweep(therm=True)
        assert o.last_acc_ratio == r.scalars()["lastAccRatio"]
        assert maxabs(o.phi[1:], r.phi()[1:]) < 1e-12
        assert maxabs(o.green[0], r.green()) < 1e-10
    assert o.rng.rand01() == r.rng_draw(1)[0]


def test_sdw_update_in_slice_live():
    p = SdwParams(L=4, m=20, s=10, rngIndex=11)
    o, r = SdwOracle(p), rb.RefSdw(p)
    # G is valid at k = m right after construction: a single-slice update there is well defined
    acc_o = o.update_in_slice(p.m)
    acc_r = r.update_in_slice(p.m)
    assert acc_o == acc_r
    assert maxabs(o.phi[1:], r.phi()[1:]) < 1e-13
    assert maxabs(o.green[0], r.green()) < 1e-11


def test_hubbard_live():
    p = HubbardParams(L=4, m=20, s=5, U=5.0, rngIndex=2)
    o, r = HubbardOracle(p), rb.RefHubbard(p)
    for sw in range(3):
        o.sweep()
        r.sweep()
        assert np.array_equal(o.aux[1:], r.aux()[1:])
        for gc in (0, 1):
            assert maxabs(o.green[gc], r.green(gc)) < 1e-9
