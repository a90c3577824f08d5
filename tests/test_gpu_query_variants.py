"""the query variants that are fan-outs of the point / line kernels, GPU vs oracle.
Statuses must be identical; probabilities (double, computed by the host libm on both sides from
bit-identical float log-odds) must be bit-identical too."""
import numpy as np
import pytest

import oracle
import volumetric_mapping_b200 as vm

pytestmark = pytest.mark.gpu
I4 = np.eye(4)


@pytest.fixture(scope="module", params=[1, 0], ids=["unknown_is_occupied", "unknown_is_unknown"])
def worlds(request):
    rng = np.random.default_rng(31)
    kw = dict(resolution=0.2, sensor_max_range=6.0, treat_unknown_as_occupied=request.param)
    gw, ow = vm.OctomapWorld(**kw), oracle.OracleWorld(**kw)
    for _ in range(3):
        d = rng.normal(size=(5000, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * rng.uniform(2.0, 5.0, (5000, 1))).astype(np.float32)
        gw.insertPointcloud(I4, pts)
        ow.insertPointcloud(I4, pts)
    return gw, ow, rng


def test_true_status_and_probability(worlds):
    gw, ow, rng = worlds
    q = rng.uniform(-6, 6, (40000, 3))
    q[:5] = [[1e9, 0, 0]] * 5
    gs, gp = gw.getCellProbabilityPoint(q)
    os_, op = ow.getCellProbabilityPoint(q)
    assert np.array_equal(gs, os_)
    assert np.array_equal(gp.view(np.uint64), op.view(np.uint64))
    assert set(np.unique(gs).tolist()) == {0, 1, 2}
    assert np.all(gp[gs == 2] == -1.0) and np.all((gp[gs != 2] > 0) & (gp[gs != 2] < 1))
    assert np.array_equal(gw.getCellTrueStatusPoint(q), os_)


@pytest.mark.parametrize("stop", [True, False])
def test_visibility(worlds, stop):
    gw, ow, rng = worlds
    q = np.concatenate([rng.uniform(-1, 1, (30000, 3)), rng.uniform(-5.5, 5.5, (30000, 3))], axis=1)
    g, o = gw.getVisibility(q, stop), ow.getVisibility(q, stop)
    assert np.array_equal(g, o)
    assert len(set(g.tolist())) >= 2


@pytest.mark.parametrize("box", [(0.5, 0.5, 0.3), (0.2, 0.0, 0.45), (0.05, 0.05, 0.05)])
def test_line_status_bounding_box(worlds, box):
    gw, ow, rng = worlds
    q = np.concatenate([rng.uniform(-1.5, 1.5, (3000, 3)), rng.uniform(-4, 4, (3000, 3))], axis=1)
    g, o = gw.getLineStatusBoundingBox(q, box), ow.getLineStatusBoundingBox(q, box)
    assert np.array_equal(g, o)
    assert len(set(g.tolist())) >= 2
