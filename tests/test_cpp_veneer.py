"""The C++ host veneer (neon_b200/cpp: Neon:: names over the C ABI) and the lid-driven-cavity benchmark built on it.

CPU part: it builds, fails loudly without a device, rejects bad command lines.
GPU part (-m gpu): the benchmark binary — the reference's own flow: host forEachActiveCell set-up, updateDeviceData, flag
halo update, computeWallNghMask, LbmIterationD3Q19 over Skeletons — reproduces the golden dumps of the UNMODIFIED
reference bit for bit in REFERENCE arithmetic, on one partition and on 2-4 partitions with every OCC / transfer mode /
halo semantic; D3Q27 against the oracle; the --visual profiles against the known answers of the stock reference run
(SURVEY.md §8c).
"""
import glob
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "neon_b200", "cpp")
APP = os.path.join(CPP, "bin", "lbm-lid-driven-cavity-flow")


@pytest.fixture(scope="module")
def app():
    if not os.path.exists(os.path.join(ROOT, "neon_b200", "lib", "libneon_lbm.so")):
        from neon_b200 import build as B
        B.build()
    subprocess.check_call(["make", "-s", "-C", CPP])
    assert os.path.exists(APP)
    return APP


def run_app(app, args, cwd, check=True):
    r = subprocess.run([app] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=600)
    if check:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


# ------------------------------------------------------------------------------------------------------------ CPU
def test_builds_and_links_the_kernel_library(app):
    out = subprocess.run(["ldd", app], capture_output=True, text=True).stdout
    assert "libneon_lbm.so" in out and "not found" not in out.split("libneon_lbm.so")[1].split("\n")[0]


def test_bad_command_line_prints_synopsis(app, tmp_path):
    r = run_app(app, ["--deviceType", "gpu"], tmp_path, check=False)
    assert r.returncode != 0 and "SYNOPSIS" in r.stdout
    r = run_app(app, ["--deviceType", "gpu", "--deviceIds", "0", "--frobnicate"], tmp_path, check=False)
    assert r.returncode != 0 and "unknown option --frobnicate" in r.stdout


def test_help_prints_synopsis_and_succeeds(app, tmp_path):
    r = run_app(app, ["--help"], tmp_path, check=False)
    assert r.returncode == 0 and "SYNOPSIS" in r.stdout and "--grid <dGrid|bGrid|eGrid>" in r.stdout


def test_cpu_device_type_is_refused(app, tmp_path):
    """north star: no CPU fallback — the reference's CPU numbers come from the reference itself"""
    r = run_app(app, ["--deviceType", "cpu", "--deviceIds", "0", "--domain-size", "16", "--max-iter", "2", "--benchmark"], tmp_path,
                check=False)
    assert r.returncode == 1 and "no CPU compute path" in r.stderr


def test_fails_loudly_without_a_device(app, tmp_path):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_app(app, ["--deviceType", "gpu", "--deviceIds", "0", "--domain-size", "16", "--max-iter", "2", "--benchmark"], tmp_path,
                check=False)
    assert r.returncode == 1 and "NeonException" in r.stderr and "no CPU fallback" in r.stderr


def test_host_logic_of_the_veneer(app):
    """partition rule, spans per data view, Skeleton schedules (Occ none/standard), Loader tokens, CellType codec, bGrid blocks and
    ghost layers, report writer, refusal to compute without Runtime::stream — neon_b200/cpp/apps/host-logic (no GPU needed)"""
    r = subprocess.run([os.path.join(CPP, "bin", "host-logic")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr


def test_sweep_driver_enumerates_the_reference_matrix():
    """sweep.py reproduces the reference sweep's axes (lbm-lid-driven-cavity-flow.py:1-10) minus cpu / eGrid"""
    sweep = os.path.join(CPP, "apps", "lbm-lid-driven-cavity-flow", "sweep.py")
    out = subprocess.run(["python", sweep, "--dry-run", "--gpus", "2"], capture_output=True, text=True, timeout=60).stdout
    assert "'n': 512" in out and "'grid': 'bGrid'" in out and "'devs': '0 1'" in out
    assert "'store': 'double', 'compute': 'float'" not in out
    # sizes x (dGrid and eGrid: 3 precision pairs x 2 device sets each + bGrid: 2 pairs x 2)
    assert int(out.strip().splitlines()[-1].split()[0]) == 8 * (3 * 2 + 3 * 2 + 2 * 2)


# ------------------------------------------------------------------------------------------------------------ GPU
GOLDEN = [  # name, dim, fp, geom
    ("cavity16_f32", (16, 16, 16), "float", "cavity"),
    ("sphere16_f64", (16, 16, 16), "double", "sphere"),
    ("sphere24_f32", (24, 24, 24), "float", "sphere"),
    ("sphere20x12x16_f32", (20, 12, 16), "float", "sphere"),
    ("cavity12_f64", (12, 12, 12), "double", "cavity"),
]


def _dump(app, tmp_path, dim, fp, geom, iters, extra=(), devices=(0,), arith="reference", lattice="D3Q19"):
    out = os.path.join(tmp_path, "dump.bin")
    args = ["--deviceType", "gpu", "--deviceIds", *devices, "--grid", "dGrid", "--dim", *dim, "--max-iter", iters, "--warmup-iter", 0,
            "--computeFP", fp, "--storageFP", fp, "--benchmark", "--geom", geom, "--arith", arith, "--lattice", lattice, "--dump", out,
            "--report-filename", os.path.join(tmp_path, "report"), *extra]
    run_app(app, args, tmp_path)
    from oracle import oracle as O
    return O.read_ref_dump(out)


@pytest.mark.gpu
@pytest.mark.parametrize("name,dim,fp,geom", GOLDEN)
def test_benchmark_flow_reproduces_the_reference_dumps(app, tmp_path, golden_dir, name, dim, fp, geom):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    d = _dump(app, str(tmp_path), dim, fp, geom, int(g["iters"]))
    assert abs(d["omega"] - float(g["omega"])) == 0.0
    assert np.array_equal(d["cls"], g["cls"]) and np.array_equal(d["mask"], g["mask"]), "flags / wall masks must be bit-exact"
    assert np.array_equal(d["pop"].view(np.uint8), g["pop"].view(np.uint8)), "REFERENCE arithmetic must be bit-exact"
    # the default (FAST) arithmetic stays within the north-star tolerance
    f = _dump(app, str(tmp_path), dim, fp, geom, int(g["iters"]), arith="fast")
    tol = 1e-5 if fp == "float" else 1e-12
    assert np.abs(f["pop"].astype(np.float64) - g["pop"]).max() / np.abs(g["pop"]).max() < tol
    rep = glob.glob(os.path.join(str(tmp_path), "report_*.json"))
    assert rep, "no report written"
    j = json.load(open(rep[0]))
    for key in ("Re", "ulb", "N", "omega", "occ", "transferMode", "transferSemantic", "MLUPS", "Loop Time (microseconds)",
                "Problem Setup Time (microseconds)", "Neon Grid Init Time (microseconds)", "Backend"):
        assert key in j, key


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [2, 3, 4])
@pytest.mark.parametrize("occ", ["--nOCC", "--sOCC"])
@pytest.mark.parametrize("mode", ["--get", "--put"])
@pytest.mark.parametrize("sem", ["--huLattice", "--huGrid"])
def test_partitions_match_the_single_partition_reference(app, tmp_path, golden_dir, parts, occ, mode, sem):
    """several z-slab partitions (an oversubscribed device list, as the reference's domain tests use) == 1 partition"""
    g = np.load(os.path.join(golden_dir, "sphere24_f32.npz"))
    d = _dump(app, str(tmp_path), (24, 24, 24), "float", "sphere", int(g["iters"]), extra=(occ, mode, sem), devices=(0,) * parts)
    assert np.array_equal(d["mask"], g["mask"])
    assert np.array_equal(d["pop"].view(np.uint8), g["pop"].view(np.uint8))


@pytest.mark.gpu
def test_ragged_partitions_and_device_setup(app, tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "sphere20x12x16_f32.npz"))
    for extra in (("--sOCC",), ("--sOCC", "--device-setup"), ("--nOCC", "--device-setup", "--put")):
        d = _dump(app, str(tmp_path), (20, 12, 16), "float", "sphere", int(g["iters"]), extra=extra, devices=(0, 0, 0))
        assert np.array_equal(d["pop"].view(np.uint8), g["pop"].view(np.uint8)), extra
    d = _dump(app, str(tmp_path), (20, 12, 16), "float", "sphere", int(g["iters"]), extra=("--graph",))
    assert np.array_equal(d["pop"].view(np.uint8), g["pop"].view(np.uint8)), "CUDA-graph replay"


@pytest.mark.gpu
@pytest.mark.parametrize("dim,warm", [((100, 80, 60), 0), ((128, 64, 64), 7)])
def test_small_box_launch_chain_matches_the_oracle(app, tmp_path, oracle, dim, warm):
    """One dense partition of more than ~400 000 cells in benchmark mode: the app runs up to 10 iterations per library call
    (LbmIterationT::runMany -> nlbm_dense_step_n, the chain of dependent launches) — the oracle's bits, with a chunk that ends at
    the warm-up boundary and an odd remainder."""
    O = oracle
    iters = 23
    cls = O.classify(O.GEOM_CAVITY_SPHERE, *dim)
    mask = O.wall_mask(19, cls)
    ref = O.run(19, O.init_pop(19, cls, np.float32), cls, mask, O.omega_cavity(dim[0]), iters)  # (the app's N is dim.x)
    d = _dump(app, str(tmp_path), dim, "float", "sphere", iters, extra=("--warmup-iter", warm))
    assert d["omega"] == O.omega_cavity(dim[0])
    assert np.array_equal(d["mask"], mask)
    assert np.array_equal(d["pop"].view(np.uint8), ref.view(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("fp,parts", [("double", 1), ("double", 3), ("float", 2)])
def test_d3q27_against_the_oracle(app, tmp_path, oracle, fp, parts):
    O = oracle
    n, iters = 20, 12
    dt = np.float64 if fp == "double" else np.float32
    cls = O.classify(O.GEOM_CAVITY_SPHERE, n, n, n)
    mask = O.wall_mask(27, cls)
    ref = O.run(27, O.init_pop(27, cls, dt), cls, mask, O.omega_cavity(n), iters)
    d = _dump(app, str(tmp_path), (n, n, n), fp, "sphere", iters, extra=("--sOCC",), devices=(0,) * parts, lattice="D3Q27")
    assert np.array_equal(d["mask"], mask)
    assert np.array_equal(d["pop"].view(np.uint8), ref.view(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("name,fp,geom,grid", [("d3q27_sphere16_f64_cfgomega", "double", "sphere", "dGrid"),
                                               ("d3q27_cavity16_f32_cfgomega", "float", "cavity", "dGrid"),
                                               ("d3q27_sphere16_f64_cfgomega", "double", "sphere", "bGrid")])
def test_d3q27_against_reference_compiled_dumps(app, tmp_path, golden_dir, name, fp, geom, grid):
    """D3Q27 through the C++ benchmark app against dumps of the reference's OWN D3Q27 kernels (apps/lbmMultiRes stream /
    collideBGK, compiled unmodified by oracle/Makefile.ref27), omega as Config.cpp:105-111 computes it for N = 16."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    d = _dump(app, str(tmp_path), (16, 16, 16), fp, geom, int(g["iters"]), extra=("--grid", grid), lattice="D3Q27")
    if fp == "double":
        assert d["omega"] == float(g["omega"])
    assert np.array_equal(d["cls"], g["cls"]) and np.array_equal(d["mask"], g["mask"])
    assert np.array_equal(d["pop"].view(np.uint8), g["pop"].view(np.uint8))


@pytest.mark.gpu
def test_visual_mode_profiles_match_the_stock_reference_run(app, tmp_path):
    """Known answers of the UNMODIFIED reference's --visual run, N=64 fp32, state after 100 iterations."""
    run_app(app, ["--deviceType", "gpu", "--deviceIds", "0", "--domain-size", "64", "--max-iter", "101", "--computeFP", "float",
                  "--storageFP", "float", "--visual", "--arith", "reference", "--report-filename", os.path.join(str(tmp_path), "r")],
            str(tmp_path))
    y = np.loadtxt(os.path.join(str(tmp_path), "NeonUniformLBM_00100_Y.dat"))
    x = np.loadtxt(os.path.join(str(tmp_path), "NeonUniformLBM_00100_X.dat"))
    known_y = {0.125: -2.29027e-05, 0.25: -0.000560718, 0.5: -0.00134673, 0.75: -0.00235457, 0.875: -0.00266163,
               0.96875: 0.0323317, 0.984375: 0.04}
    known_x = {0.125: 0.0013653, 0.5: -2.50035e-05, 0.875: -0.00140206}
    for tab, known in ((y, known_y), (x, known_x)):
        for pos, val in known.items():
            row = tab[np.argmin(np.abs(tab[:, 0] - pos))]
            assert abs(row[0] - pos) < 1e-9
            assert abs(row[1] - val) <= 1e-6 * max(1e-3, abs(val)) + 5e-6 * abs(val), (pos, row[1], val)
    assert os.path.exists(os.path.join(str(tmp_path), "u_00100.vtk")) and os.path.exists(os.path.join(str(tmp_path), "rho_00000.vtk"))


@pytest.mark.gpu
def test_visual_mode_profiles_fp64_match_the_stock_reference_run(app, tmp_path):
    """The second set of known answers of the unmodified reference: N=32, fp64/fp64, iteration 100, u_x(y)."""
    run_app(app, ["--deviceType", "gpu", "--deviceIds", "0", "0", "--domain-size", "32", "--max-iter", "101", "--computeFP", "double",
                  "--storageFP", "double", "--visual", "--sOCC", "--arith", "reference", "--report-filename", os.path.join(str(tmp_path), "r")],
            str(tmp_path))
    y = np.loadtxt(os.path.join(str(tmp_path), "NeonUniformLBM_00100_Y.dat"))
    for pos, val in {0.25: -0.000756306, 0.5: -0.00128716, 0.75: -0.00239911, 0.9375: 0.0294433}.items():
        row = y[np.argmin(np.abs(y[:, 0] - pos))]
        assert abs(row[0] - pos) < 1e-9 and abs(row[1] - val) <= 6e-6 * abs(val), (pos, row[1], val)


# ------------------------------------------------------------------------------------------------------------ bGrid
def _dump_grid(app, tmp_path, grid, *a, **k):
    extra = tuple(k.pop("extra", ())) + ("--grid", grid)
    return _dump(app, tmp_path, *a, extra=extra, **k)


@pytest.mark.gpu
@pytest.mark.parametrize("name,dim,fp,geom", GOLDEN)
def test_bgrid_reproduces_the_reference_dumps(app, tmp_path, golden_dir, name, dim, fp, geom):
    """bGrid (8^3 blocks) gives the same bytes as dGrid — as it does in the reference — including
    boxes that are not a multiple of the block edge."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    d = _dump_grid(app, str(tmp_path), "bGrid", dim, fp, geom, int(g["iters"]))
    assert np.array_equal(d["cls"], g["cls"]) and np.array_equal(d["mask"], g["mask"])
    assert np.array_equal(d["pop"].view(np.uint8), g["pop"].view(np.uint8))
    d = _dump_grid(app, str(tmp_path), "bGrid", dim, fp, geom, int(g["iters"]), extra=("--device-setup",))
    assert np.array_equal(d["mask"], g["mask"])
    assert np.array_equal(d["pop"].view(np.uint8), g["pop"].view(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("parts,occ,mode,sem", [(2, "--sOCC", "--get", "--huLattice"), (2, "--nOCC", "--put", "--huGrid"),
                                                (3, "--sOCC", "--put", "--huLattice"), (3, "--nOCC", "--get", "--huLattice")])
def test_bgrid_partitions_match_the_oracle(app, tmp_path, oracle, parts, occ, mode, sem):
    """bGrid over several block-layer partitions with a per-cardinality halo (the reference's is NaN for Q = 19)"""
    O = oracle
    dim, iters = (24, 16, 48), 12  # 6 block layers
    cls = O.classify(O.GEOM_CAVITY_SPHERE, *dim)
    mask = O.wall_mask(19, cls)
    ref = O.run(19, O.init_pop(19, cls, np.float32), cls, mask, O.omega_cavity(dim[0]), iters)
    d = _dump_grid(app, str(tmp_path), "bGrid", dim, "float", "sphere", iters, extra=(occ, mode, sem), devices=(0,) * parts)
    assert np.array_equal(d["mask"], mask)
    assert np.array_equal(d["pop"].view(np.uint8), ref.view(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("grid,extra", [("dGrid", ("--sOCC", "--get")), ("dGrid", ("--nOCC", "--put")), ("dGrid", ("--sOCC", "--put", "--huGrid")),
                                        ("bGrid", ("--sOCC", "--get")), ("bGrid", ("--nOCC", "--put"))])
def test_two_real_devices_one_process(app, tmp_path, oracle, grid, extra):
    """the reference's process model: ONE process drives both GPUs; faces move as peer stores/loads over NVLink, ordered by
    CUDA events across the devices"""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    O = oracle
    dim, iters = (40, 24, 32), 25
    cls = O.classify(O.GEOM_CAVITY_SPHERE, *dim)
    mask = O.wall_mask(19, cls)
    ref = O.run(19, O.init_pop(19, cls, np.float32), cls, mask, O.omega_cavity(dim[0]), iters)
    d = _dump_grid(app, str(tmp_path), grid, dim, "float", "sphere", iters, extra=extra, devices=(0, 1))
    assert np.array_equal(d["mask"], mask)
    assert np.array_equal(d["pop"].view(np.uint8), ref.view(np.uint8))


# ------------------------------------------------------------------------------------------- generic lambda containers
GEN = os.path.join(CPP, "bin", "generic-containers")


def test_generic_container_app_builds_with_the_generic_kernel(app):
    """Grid::newContainer(name, loadingLambda) instantiates the generic span kernel for every user lambda (nvcc)"""
    assert os.path.exists(GEN)
    sass = subprocess.run(["cuobjdump", "-sass", GEN], capture_output=True, text=True).stdout
    assert sass.count("neonLambdaOnSpan") >= 3 and sass.count("neonLambdaOnBlocks") >= 2 and "sm_100a" in sass


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [(0,), (0, 0), (0, 0, 0, 0), (0, 1)])
def test_generic_lambda_containers(app, tmp_path, devices):
    """user-written MAP / STENCIL / LBM device lambdas through dGrid::newContainer and bGrid::newContainer + Skeleton (every
    Occ, both transfer modes, halo updates; a map -> stencil -> map sequence in one Skeleton) on 1-4 partitions"""
    torch = pytest.importorskip("torch")
    if max(devices) >= torch.cuda.device_count():
        pytest.skip("needs two GPUs")
    r = subprocess.run([GEN, "--deviceIds", *[str(d) for d in devices], "--n", "36"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith(("PASS", "FAIL"))]
    # 2 axpy (dGrid, sparse bGrid) + 4 Occ x (2 modes x 2 grids diffusion + sequence + LBM)
    assert len(lines) == 2 + 4 * (4 + 1 + 1) and all(l.startswith("PASS") for l in lines), r.stdout
