"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares; the
reference ABI structs have the reference's exact layout; option defaults agree with the oracle."""
import ctypes
import os
import re
import subprocess

from forces_resilient_planner_b200 import _lib, forces
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in ("nmpc_b200.h", "FORCESNLPsolver_normal.h", "FORCESNLPsolver_final.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b((?:nmpc_|FORCESNLPsolver_(?:normal|final)_solve)\w*)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol(cuda_lib):
    declared = _declared_symbols()
    assert {"nmpc_solve_batch_f64", "FORCESNLPsolver_normal_solve", "FORCESNLPsolver_final_solve",
            "nmpc_kkt_backsolve_f64", "nmpc_pack_params_f64"} <= declared
    missing = [s for s in sorted(declared) if not hasattr(cuda_lib, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared


def test_reference_struct_layouts():
    assert ctypes.sizeof(forces.ForcesParams) == 23600
    assert forces.ForcesParams.xinit.offset == 0 and forces.ForcesParams.x0.offset == 72
    assert forces.ForcesParams.all_parameters.offset == 2792 and forces.ForcesParams.num_of_threads.offset == 23592
    assert ctypes.sizeof(forces.ForcesOutput) == 2720 and forces.ForcesOutput.x20.offset == 19 * 136
    info = forces.ForcesInfo
    assert ctypes.sizeof(info) == 136
    names = ("it", "it2opt", "res_eq", "res_ineq", "rsnorm", "rcompnorm", "pobj", "dobj", "dgap", "rdgap", "mu",
             "mu_aff", "sigma", "lsit_aff", "lsit_cc", "step_aff", "step_cc", "solvetime", "fevalstime")
    assert [getattr(info, f).offset for f in names] == \
        [0, 4, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 100, 104, 112, 120, 128]


def test_headers_compile_as_c_and_cpp(tmp_path):
    src = '#include "FORCESNLPsolver_normal.h"\n#include "FORCESNLPsolver_final.h"\n#include "nmpc_b200.h"\n' \
          'int main(void){FORCESNLPsolver_normal_params p; FORCESNLPsolver_final_output o; (void)p; (void)o; ' \
          'return sizeof(p)==23600 && sizeof(o)==2720 ? 0 : 1;}\n'
    for ext, cc in ((".c", "gcc"), (".cpp", "g++")):
        f = tmp_path / ("t" + ext)
        f.write_text(src)
        exe = tmp_path / ("t" + ext + ".out")
        subprocess.check_call([f"/usr/bin/{cc}", "-I", os.path.join(ROOT, "include"), str(f), "-o", str(exe)])
        assert subprocess.call([str(exe)]) == 0


def test_default_options_match_the_oracle(cuda_lib):
    a, b = _lib.default_opts(), O.default_opts()
    for name, _ in _lib.NmpcOpts._fields_:
        assert getattr(a, name) == getattr(b, name), name
    assert a.maxit == 200 and a.tol_stat == a.tol_eq == a.tol_ineq == a.tol_comp == 1e-4


def test_metadata_calls_work_without_a_gpu(cuda_lib):
    assert cuda_lib.nmpc_supported_horizon(20) == 1 and cuda_lib.nmpc_supported_horizon(40) == 1
    assert cuda_lib.nmpc_supported_horizon(21) == 0
    assert 30_000 < cuda_lib.nmpc_smem_bytes(20, 8, 8) < 60_000
    assert cuda_lib.nmpc_backsolve_factor_words() == 204
    assert b"sm_100a" in cuda_lib.nmpc_version()


def test_no_cpu_fallback_without_a_device(cuda_lib):
    """On a machine without a GPU the solve entry points must fail loudly, not compute."""
    import torch
    if torch.cuda.is_available():
        return
    from forces_resilient_planner_b200 import solver as S, workloads as W
    try:
        S.solve_host(W.config2(2))
    except RuntimeError as e:
        assert "rc=-101" in str(e)
    else:
        raise AssertionError("solve_host computed something without a GPU")
    # the reference-named symbols answer in the reference's own vocabulary (header :110-139): "no CUDA device" becomes
    # LICENSE_ERROR (-100, "solver not valid on this machine"), which the planner treats like any failed solve
    w = forces.FORCESNormal()
    assert w.solve_params() == -100


def _build_dropin_probe(tmp_path):
    """The drop-in tree (same directory / header / archive names as plan_manage/solver/) and a planner-side translation
    unit that uses only the reference's header names, macros and symbols, linked with the reference's own link line
    (plan_manage/CMakeLists.txt:58-65 include + link directories, :82-83 `libFORCESNLPsolver_normal.a
    libFORCESNLPsolver_final.a` by file name) -- no cudart, no extra library."""
    host = os.path.join(ROOT, "forces_resilient_planner_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s", "dropin"])
    d = os.path.join(ROOT, "build", "dropin", "solver")
    exe = str(tmp_path / "dropin_tu")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "tools", "dropin_tu.cpp"),
                           "-I", f"{d}/normal/FORCESNLPsolver_normal/include", "-I", f"{d}/final/FORCESNLPsolver_final/include",
                           "-L", f"{d}/normal/FORCESNLPsolver_normal/lib", "-L", f"{d}/final/FORCESNLPsolver_final/lib",
                           "-l:libFORCESNLPsolver_normal.a", "-l:libFORCESNLPsolver_final.a"])
    return exe


def test_dropin_archives_link_with_the_reference_link_line(cuda_lib, tmp_path):
    import torch
    exe = _build_dropin_probe(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert out.returncode == 0 and "exitflags 1 1 maxit 200" in out.stdout, out.stdout + out.stderr
    else:       # no device: both calls answer LICENSE_ERROR and say why; nothing is computed on the CPU
        assert out.returncode == 100 and "exitflags -100 -100" in out.stdout, out.stdout + out.stderr
        assert "exitflag -100" in out.stdout and "cuda" in out.stdout.lower()
    # a library that cannot be found is the same failure, reported by the stub itself
    env = dict(os.environ, NMPC_B200_LIB="/nonexistent/libnmpc_b200.so")
    host = os.path.join(ROOT, "forces_resilient_planner_b200", "host")
    exe2 = str(tmp_path / "stub_only")
    subprocess.run(["/usr/bin/gcc", "-O2", "-o", exe2, "-x", "c", "-", "-I", os.path.join(ROOT, "include"),
                    "-DNMPC_STUB_VARIANT=normal", '-DNMPC_STUB_HEADER="FORCESNLPsolver_normal.h"',
                    os.path.join(host, "forces_stub.c")], check=True, input=b"""
#include "FORCESNLPsolver_normal.h"
int main(void){ static FORCESNLPsolver_normal_params p; static FORCESNLPsolver_normal_output o; static FORCESNLPsolver_normal_info i;
  return FORCESNLPsolver_normal_solve(&p, &o, &i, 0, 0) == LICENSE_ERROR_FORCESNLPsolver_normal ? 0 : 1; }""")
    # without a default path and with a bad NMPC_B200_LIB the loader's search path is the last resort: not found here
    out = subprocess.run([exe2], capture_output=True, text=True, env=dict(env, LD_LIBRARY_PATH=""))
    assert out.returncode == 0 and "cannot load libnmpc_b200.so" in out.stderr


def test_product_never_touches_the_oracle():
    """Neither the package nor the measurement scripts may import, link or execute anything under oracle/
    (only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may)."""
    sources = []
    for top in ("forces_resilient_planner_b200", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            sources += [os.path.join(dirpath, fn) for fn in files
                        if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h", ".sh"))]
    assert len(sources) > 20
    for path in sources:
        txt = open(path).read()
        assert "import oracle" not in txt and "from oracle" not in txt, path
        assert "libnmpc_oracle" not in txt and not re.search(r'#include\s*[<"][^>"]*oracle', txt), path


def test_argument_checks_of_the_kernels_around_the_solve(cuda_lib):
    """Bad sizes / null pointers are rejected with a negative NMPC_ERR_ARG code before any CUDA call,
    so this runs without a GPU."""
    vp = ctypes.c_void_p
    one = ctypes.c_void_p(16)          # any non-null pointer: never dereferenced on these paths
    f = cuda_lib.nmpc_sample_reference_f64
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double] + [vp] * 9
    assert f(4, 20, 8, 0.0, one, one, one, one, None, one, one, None, None) < 0          # Ts must be positive
    assert f(4, 20, 8, 0.05, None, one, one, one, None, one, one, None, None) < 0        # null path
    assert f(0, 20, 8, 0.05, None, None, None, None, None, None, None, None, None) == 0  # empty batch is a no-op
    g = cuda_lib.nmpc_propagate_ellipsoids_f64
    g.restype = ctypes.c_int
    g.argtypes = [ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(_lib.EllipsoidConsts), vp, vp]
    c = _lib.EllipsoidConsts()
    cuda_lib.nmpc_default_ellipsoid_consts(ctypes.byref(c))
    assert (c.mass, c.drag, c.ego_r, c.ego_h, c.ext_noise_bound, c.epsilon, c.Ts) == (0.745319, 0.33, 0.27, 0.0425, 0.5, 0.06, 0.05)
    assert g(4, 20, None, ctypes.byref(c), one, None) < 0
    c.mass = 0.0
    assert g(4, 20, one, ctypes.byref(c), one, None) < 0                                 # bad constants
    h = cuda_lib.nmpc_select_corridors_f64
    h.restype = ctypes.c_int
    h.argtypes = [ctypes.c_int] * 5 + [vp, ctypes.c_longlong] + [vp] * 4 + [ctypes.POINTER(ctypes.c_double)] + [vp] * 7
    args = [one, 3 * 64, one, one, one, one, None, one, one, one, one, one, one, None]
    assert h(4, 20, 64, 8, 6, *args) < 0                                                 # fewer than 7 rows cannot hold the box
    assert h(4, 20, 64, 0, 30, *args) < 0
    bad = list(args); bad[1] = 10                                                        # stride shorter than a cloud
    assert h(4, 20, 64, 8, 30, *bad) < 0
    assert b"bad argument" in cuda_lib.nmpc_last_error()
    # misaligned device pointers are rejected by the entry points that move per-problem blocks with TMA
    k = cuda_lib.nmpc_kkt_backsolve_f64
    k.restype = ctypes.c_int
    k.argtypes = [ctypes.c_int, ctypes.c_int] + [vp] * 6
    assert k(2, 20, vp(16), vp(32), vp(48), vp(64), vp(72), None) < 0 and b"16-byte" in cuda_lib.nmpc_last_error()
    o = _lib.default_opts()
    sv = cuda_lib.nmpc_solve_batch_f64
    assert sv(2, 20, 6, vp(8), vp(16), vp(32), vp(40), vp(64), 0, ctypes.byref(o), vp(96), vp(128), vp(160), None) < 0
    assert b"16-byte" in cuda_lib.nmpc_last_error()


def test_argument_checks_of_the_multi_gpu_entry_points(cuda_lib):
    """nmpc_peers_* / nmpc_solve_batch_sharded_* reject bad arguments before any CUDA or NCCL call (runs without a GPU)."""
    vp, i = ctypes.c_void_p, ctypes.c_int
    mk = cuda_lib.nmpc_peers_create
    mk.restype = i
    mk.argtypes = [i, i, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(vp)]
    out = vp()
    assert mk(0, 0, 1024, 16, ctypes.byref(out)) == -11            # world < 1
    assert mk(17, 0, 1024, 16, ctypes.byref(out)) == -11           # more ranks than a node's peer table holds
    assert mk(2, 2, 1024, 16, ctypes.byref(out)) == -11            # rank out of range
    assert mk(2, 0, 0, 16, ctypes.byref(out)) == -11               # empty buffer
    assert b"world=" in cuda_lib.nmpc_last_error() and not out.value
    for name, n_extra in (("nmpc_solve_batch_sharded_p2p_f64", 2), ("nmpc_solve_batch_sharded_p2p_f32", 1)):
        f = getattr(cuda_lib, name)
        f.restype = i
        f.argtypes = [vp, i, i, i] + [vp] * 5 + [i, vp, vp] + ([i, vp] if n_extra == 2 else [vp])
        assert f(None, 4, 20, 8, None, None, None, None, None, 0, None, None, *([0, None] if n_extra == 2 else [None])) == -11
    for name in ("nmpc_peers_barrier", "nmpc_peers_status", "nmpc_peers_export", "nmpc_peers_connect"):
        f = getattr(cuda_lib, name)
        f.restype = i
        f.argtypes = [vp, vp] if name != "nmpc_peers_status" else [vp]
        assert f(*([None] * len(f.argtypes))) == -11
    cuda_lib.nmpc_peers_destroy.argtypes = [vp]
    assert cuda_lib.nmpc_peers_destroy(None) == 0
    sh = cuda_lib.nmpc_solve_batch_sharded_f64
    sh.restype = i
    sh.argtypes = [vp, i, i, i] + [vp] * 5 + [i, vp, vp, vp, vp, i, vp]
    assert sh(None, 4, 20, 8, None, None, None, None, None, 0, None, None, None, None, 0, None) < 0
