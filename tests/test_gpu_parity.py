"""-m gpu: the CUDA path (through the C ABI of libogn_b200.so) against the CPU checker on the same seeded inputs.
Bar: every CSDR, activation, weight and the ComputeSystem RNG state bit-identical (integer work and, thanks to the
order-preserving sums and the glibc expf/tanhf replay, the fp32 work too — so the 1e-5 tolerance north_star allows is
not needed)."""
import numpy as np
import pytest

from ogmaneo2_b200 import configs
from ogmaneo2_b200.flat_api import InputType, LayerDesc

from common import maker, run_pair

pytestmark = pytest.mark.gpu


def _check(product, oracle, cfg, steps, **kw):
    flat, kind = oracle
    flat.set_num_threads(1)  # the reference races on cs.rng with more threads and we compare its state
    err = run_pair(cfg, steps, maker(product.lib()), maker(flat), compare_rng=True, **kw)
    assert err is None, f"[checker={kind}] {err}"


def test_c1_sine(product, oracle):
    _check(product, oracle, configs.c1_sine(), 400)


def test_c1_contract_length_known_answer(product, oracle):
    """Config 1 at its contract length (10 000 steps, learning on): the number of mispredicted steps among the
    last 1000 equals the checker's (19 for the reference)."""
    from common import c1_last1000_errors
    flat, kind = oracle
    flat.set_num_threads(1)
    cfg = configs.c1_sine()
    want = c1_last1000_errors(maker(flat)(cfg), cfg)
    got = c1_last1000_errors(maker(product.lib())(cfg), cfg)
    assert got == want == 19, f"[checker={kind}] product {got}, checker {want}"


def test_c2_video_small(product, oracle):
    _check(product, oracle, configs.c2_video(12, 3), 60)


def test_c3_actor(product, oracle):
    _check(product, oracle, configs.c3_actor(8), 300)


def test_c4_large_layer_small(product, oracle):
    _check(product, oracle, configs.c4_large(24), 30)


def test_c5_unit(product, oracle):
    _check(product, oracle, configs.c5_unit(16), 40, check_every=5)


def _variant_counts(product):
    """launches per (kernel kind, variant) since the library was loaded: rows sc_forward, sc_recon, pred_step;
    columns scalar, vec1, vec2, pipe4, pipe6, tma"""
    import ctypes as C
    out = (C.c_longlong * 24)()
    product.cdll().ogn_debug_variant_launches(out)
    return np.array(list(out), dtype=np.int64).reshape(3, 8)


def _check_contract_size(product, oracle, cfg, steps, want_vec1):
    """BASELINE-size parity with the DEFAULT kernel selector: every CSDR, activation and weight against the
    compiled reference running on all host threads (no Actor, so the reference is thread-count independent; its shared
    RNG is not, so the RNG state is not compared here)."""
    import os
    flat, kind = oracle
    assert "OGN_KERNELS" not in os.environ, "this test pins the default selector"
    flat.set_num_threads(os.cpu_count() or 1)
    before = _variant_counts(product)
    err = run_pair(cfg, steps, maker(product.lib()), maker(flat), check_every=3, compare_rng=False)
    flat.set_num_threads(1)
    assert err is None, f"[checker={kind}] {err}"
    used = _variant_counts(product) - before
    for row, name in want_vec1:
        assert used[row, 1] > 0, f"{name} never ran as vec1 (the headline kernel) at this size: {used[row].tolist()}"


def test_c2_video_full_size(product, oracle):
    """BASELINE config 2 at its stated size: 32x32x16 input, 3 layers of 32x32x32, radius 3."""
    _check_contract_size(product, oracle, configs.c2_video(32, 3), 24, want_vec1=[])


def test_c4_large_layer_at_oracle_size_128(product, oracle):
    """BASELINE config 4 at SURVEY's oracle size 128x128 (1.3e9 weights): the SparseCoder forward runs the headline vec1
    kernel under the default selector."""
    _check_contract_size(product, oracle, configs.c4_large(128), 10, want_vec1=[(0, "sc_forward")])


def test_c4_large_layer_160_all_headline_kernels(product, oracle):
    """160x160 (2.1e9 weights, the reference needs ~25 GB): the smallest square size at which the default selector picks
    vec1 for ALL of forward, reconstruction and Predictor (more than 8192 packs per launch), i.e. exactly the kernels
    bench.py times at 512x512."""
    _check_contract_size(product, oracle, configs.c4_large(160), 8,
                         want_vec1=[(0, "sc_forward"), (1, "sc_recon"), (2, "pred_step")])


def test_c4_large_layer_224_opt_in(product, oracle):
    """The largest size the reference can construct (224x224, 4.1e9 weights, ~45 GiB of host memory, two
    minutes of reference init). Opt-in: OGN_TEST_C4_224=1."""
    import os
    if os.environ.get("OGN_TEST_C4_224") != "1":
        pytest.skip("opt-in (OGN_TEST_C4_224=1): needs ~60 GB of host memory and ~4 minutes")
    _check_contract_size(product, oracle, configs.c4_large(224), 6,
                         want_vec1=[(0, "sc_forward"), (1, "sc_recon"), (2, "pred_step")])


@pytest.mark.parametrize("shape", [
    dict(inp=(5, 7, 6), hid=(3, 4, 5), r=1),      # downsampling, non-square, Z not a power of two
    dict(inp=(2, 3, 4), hid=(6, 5, 8), r=2),      # upsampling: fields overlap and clamp everywhere
    dict(inp=(9, 4, 3), hid=(4, 9, 33), r=3),     # hidden Z > 32 (two lane chunks), radius larger than the grid
    dict(inp=(6, 6, 40), hid=(5, 5, 7), r=1),     # visible Z > 32
])
def test_odd_geometries(product, oracle, shape):
    cfg = dict(name=f"odd{shape}", inputSizes=[shape["inp"]], inputTypes=[InputType.prediction],
               layerDescs=[LayerDesc(hiddenSize=shape["hid"], ffRadius=shape["r"], pRadius=shape["r"], temporalHorizon=2),
                           LayerDesc(hiddenSize=shape["hid"], ffRadius=shape["r"], pRadius=shape["r"], temporalHorizon=2)],
               streams=["iid"], seed=99)
    _check(product, oracle, cfg, 40)


def test_learning_disabled_and_params(product, oracle):
    """learnEnabled=False steps interleaved with learning steps, non-default alpha."""
    flat, kind = oracle
    cfg = configs.c2_video(8, 2)
    from common import make_inputs, diff_snapshots
    a, b = maker(product.lib())(cfg), maker(flat)(cfg)
    for h in (a, b):
        h.setSCAlpha(0, 0.25)
        h.setPredAlpha(0, 0, 0.125)
    for t in range(30):
        ins = make_inputs(cfg, t)
        learn = (t % 3) != 1
        a.step(ins, learn)
        b.step(ins, learn)
    d = diff_snapshots(a.snapshot(True), b.snapshot(True))
    assert not d, f"[checker={kind}] {d[:3]}"


def test_actor_more_history_samples_than_one_launch_holds(product, oracle):
    """historyIters above OGN_MAX_ACTOR_SAMPLES (16): the batched Actor learn must split into several launches and still
    apply the samples in the reference's order."""
    flat, kind = oracle
    flat.set_num_threads(1)
    cfg = configs.c3_actor(8)
    from common import make_inputs, diff_snapshots, reward_at
    a, b = maker(product.lib())(cfg), maker(flat)(cfg)
    for h in (a, b):
        h.setActorParams(1, 0.01, 0.01, 0.99, 4, 21)
    action = None
    for t in range(60):
        ins = make_inputs(cfg, t, action)
        a.step(ins, True, reward_at(t), False)
        b.step(ins, True, reward_at(t), False)
        action = b.getPredictionCs(1)
        assert np.array_equal(a.getPredictionCs(1), action), f"[checker={kind}] action differs at step {t}"
    d = diff_snapshots(a.snapshot(True), b.snapshot(True))
    assert not d, f"[checker={kind}] {d[:3]}"


@pytest.mark.parametrize("env", [
    {"OGN_KERNELS": "scalar"}, {"OGN_KERNELS": "vec1"}, {"OGN_KERNELS": "vec2"}, {"OGN_KERNELS": "tma"}, {"OGN_KERNELS": "wide"},
    {"OGN_KERNELS": "wide2"}, {"OGN_KERNELS": "stage"}, {"OGN_KERNELS": "stage", "OGN_STAGE_CHUNKS": "1"}, {"OGN_KERNELS": "vec1", "OGN_SC_UPDATE": "2"}, {"OGN_SC_UPDATE": "2", "OGN_UPD2_BATCHES": "1"},
    {"OGN_SMALL_KERNEL": "1"}, {"OGN_SMALL_KERNEL": "1", "OGN_SC_UPDATE": "2"}, {"OGN_GRAPHS": "0"}, {"OGN_SIDE_STREAM": "0"}, {"OGN_SIDE_STREAM": "1", "OGN_GRAPHS": "0"},
], ids=lambda e: "-".join(f"{k[4:].lower()}={v}" for k, v in e.items()))
def test_every_kernel_variant_and_scheduling_mode_is_bit_exact(env):
    """The alternative kernels (generic one-cell-per-lane, deeper register batches, TMA-staged CSDR windows, one warp per
    pack, Predictor weight rows staged in shared memory, the full-sector update kernel) and scheduling modes (the single-launch small-hierarchy step instead of per-phase
    launches, no CUDA graphs, one stream instead of two) are selected per process: rerun a parity subset in a subprocess for
    each. The default run of this file covers per-phase launches captured into graphs with the side stream (small
    hierarchies), `wide` for small launches, `vec1` for large ones and update v1."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
           "-k", "c2_video_small or c4_large_layer_small or c1_sine or odd or c5_unit"]
    r = subprocess.run(cmd, cwd=root, env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_parameter_changes_mid_run_with_graph_replay(product, oracle):
    """Hyper-parameters and learnEnabled change while the captured step graphs are being replayed: the graph key must
    pick up every change (alpha is a kernel argument baked into a captured launch)."""
    flat, kind = oracle
    cfg = configs.c2_video(8, 3)
    from common import make_inputs, diff_snapshots
    a, b = maker(product.lib())(cfg), maker(flat)(cfg)
    for t in range(96):
        if t in (24, 48, 72):
            for h in (a, b):
                h.setSCAlpha(t // 24 % 3, 0.05 * (t // 24))
                h.setPredAlpha(0, 0, 0.25 + 0.125 * (t // 24))
        learn = not (40 <= t < 56)
        ins = make_inputs(cfg, t)
        a.step(ins, learn)
        b.step(ins, learn)
        if t % 8 == 7:
            assert np.array_equal(a.getPredictionCs(0), b.getPredictionCs(0)), f"[checker={kind}] prediction differs at step {t}"
    d = diff_snapshots(a.snapshot(True), b.snapshot(True))
    assert not d, f"[checker={kind}] {d[:3]}"


@pytest.mark.parametrize("seed", range(16))
def test_random_shapes_inner(product, oracle, seed):
    """The same seeded random geometries the CPU suite pins the oracle on (tests/random_cfgs.py), on the CUDA path.
    Runs only inside the subprocess test_random_shapes starts."""
    import os
    if os.environ.get("OGN_RUN_RANDOM_SHAPES") != "1":
        pytest.skip("runs inside the subprocess of test_random_shapes")
    from random_cfgs import random_cfg
    _check(product, oracle, random_cfg(seed), 24, check_every=4)


# A subprocess, so that a device fault on one odd shape cannot poison the CUDA context of the tests that follow. Besides the
# default selection the random geometries also run with the opt-in paths that have code of their own per shape class.
@pytest.mark.parametrize("env", [{}, {"OGN_SMALL_KERNEL": "1"}, {"OGN_KERNELS": "tma"}, {"OGN_KERNELS": "stage"}, {"OGN_KERNELS": "wide2"}],
                         ids=lambda e: "-".join(f"{k[4:].lower()}={v}" for k, v in e.items()) or "default")
def test_random_shapes(env):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-k", "random_shapes_inner"]
    r = subprocess.run(cmd, cwd=root, env=dict(os.environ, OGN_RUN_RANDOM_SHAPES="1", **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-1000:]
