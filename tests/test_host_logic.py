"""Host-side logic that needs no GPU: config loading, state-dict schema, sharding (gloo, world 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nlml_hpe_b200 import NLML_HPE_Model_Builder as MB
from nlml_hpe_b200 import config, sharding, synthetic


def test_load_config_tolerates_equals_line(tmp_path):
    """configs/config_NLML_HPE_Test.yaml:29 of the reference is `val_set_path = "..."`."""
    p = tmp_path / "c.yaml"
    p.write_text('yaw_intervals:\n  - [-51, -33.33]\n  - [0, 16.67]\nval_set : "facescape" # x\n'
                 '# val_set_path = "commented"\nval_set_path = "E:/some/path(jpg)" \n')
    cfg = config.load_config(str(p))
    assert cfg["val_set"] == "facescape"
    assert cfg["val_set_path"] == "E:/some/path(jpg)"
    assert cfg["yaw_intervals"][0] == [-51, -33.33]
    good = tmp_path / "g.yaml"
    good.write_text("input_size: 1404\nyaw_bins:\n  min_bin: -50\n")
    assert config.load_config(str(good))["input_size"] == 1404


def test_state_dict_schema_matches_reference(state_dicts):
    """Keys/shapes of LandmarkEncoder / AnglePredictionNetwork state_dicts."""
    enc = MB.LandmarkEncoder(1404, [(1, 3)] * 3)
    assert list(enc.state_dict().keys()) == [f"encoder.{i}.{p}" for i in (0, 2, 4, 6, 8, 10) for p in ("weight", "bias")]
    shapes = [tuple(v.shape) for k, v in enc.state_dict().items() if k.endswith("weight")]
    assert shapes == [(1024, 1404), (512, 1024), (256, 512), (128, 256), (64, 128), (9, 64)]
    head = MB.AnglePredictionNetwork(3)
    assert [tuple(v.shape) for k, v in head.state_dict().items() if k.endswith("weight")] == \
        [(128, 3), (256, 128), (128, 256), (64, 128), (1, 64)]
    model = MB.build_combined_model(*state_dicts)      # shipped heads + synthetic encoder load cleanly
    assert sum(p.numel() for p in model.parameters()) == 2136585 + 3 * 74753
    with pytest.raises(Exception):
        enc(torch.zeros(1, 1404))                       # stand-alone forward is not on the hot path


def test_combined_checkpoint_roundtrip(tmp_path, state_dicts):
    model = MB.build_combined_model(*state_dicts)
    path = str(tmp_path / "combined.pth")
    torch.save(model.state_dict(), path)
    again = MB.load_combined_model(path)
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), again.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_synthetic_inputs_are_deterministic(art, rows):
    a = synthetic.make_features(64, art["W"], *rows, U_id=art["U_id"], seed=5)
    b = synthetic.make_features(64, art["W"], *rows, U_id=art["U_id"], seed=5)
    assert np.array_equal(a, b) and a.dtype == np.float32 and a.shape == (64, 1404)
    assert 1.0 < np.linalg.norm(a, axis=1).mean() < 5.0


def test_cosine_rows_reproduce_factor_matrices(art):
    """Known-answer fact 1 of SURVEY.md section 4."""
    for name, lim, tol in (("yaw", 50, 0.03), ("pitch", 40, 0.15), ("roll", 30, 0.03)):
        U = art[f"U_{name}"]
        bins = np.radians(np.linspace(-lim, lim, U.shape[0]))
        c = synthetic.cos_rows(bins, art[f"optimized_{name}"])
        assert np.abs(c - U).max() < tol


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    out = sharding.run_sharded(lambda x: x * 2.0, lambda lo, hi: full[lo:hi], n)
    q.put((rank, torch.equal(out, full * 2.0)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 10])
def test_sharded_run_gathers_in_order_gloo_world2(n):
    """N>1 path: each rank works on its slice, results are gathered once at the end (no compute collective)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + n
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(results) == [(0, True), (1, True)]


def test_plan_cache_key_follows_the_constants(art, rows):
    """TD_Tester re-uses device plans across the reference's per-sample calls; the key must change with any change of
    W, of the cosine rows or of the device, and only then."""
    from nlml_hpe_b200 import TD_Tester
    W = np.ascontiguousarray(art["W"], dtype=np.float32)
    r = [np.ascontiguousarray(x, dtype=np.float64) for x in rows]
    k0 = TD_Tester._content_key(W, r, None)
    assert TD_Tester._content_key(W.copy(), [x.copy() for x in r], None) == k0
    W2 = W.copy()
    W2[3, 1, 2, 0, 777] = np.nextafter(W2[3, 1, 2, 0, 777], np.float32(np.inf))       # one ulp in one entry
    assert TD_Tester._content_key(W2, r, None) != k0
    W3 = W.copy()
    W3[0, 0, 0, 0, 0], W3[0, 0, 0, 0, 1] = W[0, 0, 0, 0, 1], W[0, 0, 0, 0, 0]         # a swap keeps sum and xor of 32-bit words...
    assert TD_Tester._content_key(W3, r, None) != k0                                   # ...but not of the 64-bit ones
    W4 = W.copy()
    W4[[0, 1]] = W4[[1, 0]]                                                             # two identity slices swapped: same multiset of words
    assert TD_Tester._content_key(W4, r, None) != k0
    r2 = [x.copy() for x in r]
    r2[1][2, 3] += 1e-12
    assert TD_Tester._content_key(W, r2, None) != k0
    assert TD_Tester._content_key(W, r, "cuda:1") != k0
    assert TD_Tester._content_key(W.reshape(5, 3, 3, 1, 3 * 1404), r, None) != k0      # same bytes, other ranks
    # the per-object fast path: the same array object is hashed once, and an in-place change of it is noticed
    assert TD_Tester._content_key(W, r, None) == k0
    keep = W[2, 1, 1, 1, 5]
    W[2, 1, 1, 1, 5] = np.nextafter(keep, np.float32(np.inf))
    assert TD_Tester._content_key(W, r, None) != k0
    W[2, 1, 1, 1, 5] = keep
    assert TD_Tester._content_key(W, r, None) == k0
