"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from cases import conf2d
from numericalflowiteration_b200 import (Config1D, Config2D, Config3D, CudaError, CudaScheduler, F0, _lib, n_nodes, n_quad,
                                         partition, stride_t)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nufi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nufi_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(_lib.SYMBOLS) == declared
    L = _lib.load()
    for s in declared:
        assert hasattr(L, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (nufi_b200_\w+)", nm))
    assert exported == set(declared)
    assert b"sm_100a" in L.nufi_b200_version()


def test_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "oracle" not in out and "nufi_ref" not in out
    nm = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "orc_" not in nm and "ref_rho" not in nm


def test_config_struct_layout_matches_header():
    """ctypes mirrors == the C structs of include/nufi_b200.h (compiled with gcc) == config_t<double> layout."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "nufi_b200.h"
int main(void){
 printf("%zu %zu %zu %zu\n", sizeof(nufi_b200_config1d), offsetof(nufi_b200_config1d, dt), offsetof(nufi_b200_config1d, du), sizeof(nufi_b200_f0));
 printf("%zu %zu %zu\n", sizeof(nufi_b200_config2d), offsetof(nufi_b200_config2d, dt), offsetof(nufi_b200_config2d, dv));
 printf("%zu %zu %zu\n", sizeof(nufi_b200_config3d), offsetof(nufi_b200_config3d, dt), offsetof(nufi_b200_config3d, dw));
 return 0; }'''
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    want = [C.sizeof(Config1D), Config1D.dt.offset, Config1D.du.offset, C.sizeof(F0),
            C.sizeof(Config2D), Config2D.dt.offset, Config2D.dv.offset,
            C.sizeof(Config3D), Config3D.dt.offset, Config3D.dw.offset]
    assert got == want
    assert C.sizeof(Config1D) == 13 * 8 and C.sizeof(Config3D) == 35 * 8


def test_sizes():
    c1, c2, c3 = Config1D(), Config2D(), Config3D()
    assert (stride_t(c1), stride_t(c2), stride_t(c3)) == (259, 1225, 1331)
    assert (n_quad(c1), n_quad(c2), n_quad(c3)) == (131072, 16777216, 262144)
    assert (n_nodes(c1), n_nodes(c2), n_nodes(c3)) == (256, 1024, 512)
    assert (c1.Nt, c2.Nt, c3.Nt) == (1600, 800, 50)


def test_partition_is_the_reference_rule():
    """nufi/cuda_scheduler.hpp:88-111 / bin/test_nufi_gpu_3d.cpp:80-105: contiguous, first `rem` parts one longer."""
    for total, parts in ((10, 3), (262144, 8), (7, 8), (0, 4), (131072, 5)):
        edges = [partition(total, parts, p) for p in range(parts)]
        assert edges[0][0] == 0 and edges[-1][1] == total
        for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
            assert b == c
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert partition(10, 3, 1, begin=100) == (104, 107)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    with pytest.raises(CudaError) as e:
        CudaScheduler(conf2d(), F0(0, 0.05, 0.5))
    assert "no CPU fallback" in str(e.value)
    from numericalflowiteration_b200 import measure_fp64_peak

    with pytest.raises(CudaError):
        measure_fp64_peak()


def test_argument_errors_do_not_need_a_gpu():
    for bad_order in (2, 9):  # the reference instantiates orders 3..8 (nufi/cuda_kernel.cu:191-203)
        with pytest.raises(ValueError):
            CudaScheduler(conf2d(), F0(0, 0.05, 0.5), order=bad_order)
    with pytest.raises(ValueError):
        CudaScheduler(conf2d(Nx=2), F0(0, 0.05, 0.5))
    with pytest.raises(ValueError):  # grid smaller than the spline window
        CudaScheduler(conf2d(Nx=5), F0(0, 0.05, 0.5), order=6)
    with pytest.raises(ValueError, match="not periodic"):  # f0 must have the period of the box (k L = 2 pi m)
        CudaScheduler(conf2d(), F0(0, 0.05, 0.37))


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["NUFI_ROOT"]); sys.path.insert(0, os.path.join(os.environ["NUFI_ROOT"], "tests"))
from cases import CASES
from numericalflowiteration_b200 import n_quad, partition
from numericalflowiteration_b200.distributed import host_allreduce_partials
from oracle.oracle_py import Oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mk, f0 = CASES["2d-landau"]
conf = mk()
orc = Oracle()
coeffs, _, _ = orc.run(conf, f0, 6)
lo, hi = partition(n_quad(conf), world, rank)
part = orc.rho_partial(conf, f0, 6, coeffs, lo, hi)     # stands in for the device work of this rank's shard
total = host_allreduce_partials(dist, part)
want = orc.rho(conf, f0, 6, coeffs)
err = float(np.max(np.abs(1 + total - want)) / np.max(np.abs(want)))
assert err <= 1e-13, err
# every rank must hold the identical reduced vector (the replicated field tail depends on it)
g = [None] * world
dist.all_gather_object(g, total.tobytes())
assert all(x == g[0] for x in g)
print(f"rank {rank}/{world} q=[{lo},{hi}) err={err:.1e}")
# the fused multi-GPU step shards by interleaved velocity nodes instead (csrc/peer.cu): rank r takes nodes r, r+world, ...
from numericalflowiteration_b200 import n_nodes, n_vel
from numericalflowiteration_b200.distributed import velocity_share
nv, nn = n_vel(conf), n_nodes(conf)
mine = velocity_share(nv, world, rank)
shares = [None] * world
dist.all_gather_object(shares, list(mine))
assert sorted(j for sh in shares for j in sh) == list(range(nv))          # disjoint, complete
part2 = np.zeros(nn)
for l in range(0, nn, 7):                                                    # a sample of the nodes keeps this quick
    for j in mine:
        part2 = orc.rho_partial(conf, f0, 6, coeffs, l * nv + j, l * nv + j + 1, rho=part2)
total2 = host_allreduce_partials(dist, part2)
err2 = float(np.max(np.abs(1 + total2[::7] - want[::7])) / np.max(np.abs(want)))
assert err2 <= 1e-13, err2
print(f"rank {rank}/{world} velocity share {mine} err={err2:.1e}")
dist.destroy_process_group()
'''


def test_sharded_rho_world_size_2_gloo(tmp_path):
    """N > 1 host logic on CPU: reference partition of the flat q range per rank, all-reduce (gloo) of the partial rho,
    identical result on every rank.  The oracle stands in for each rank's device work."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, NUFI_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0/2" in r.stdout and "rank 1/2" in r.stdout


def test_cpp_drivers_build_and_fail_loudly_without_gpu():
    """The C++ host side (include/nufi/*.hpp + bin/*.cpp) compiles with plain g++ against the C ABI; without a GPU the
    drivers exit non-zero with the CUDA error (no silent CPU path)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "bin")], check=True)
    for name in ("test_nufi_gpu_1d", "test_nufi_gpu_2d", "test_nufi_gpu_3d", "test_nufi_cpu_1d", "test_nufi_cpu_2d", "test_nufi_cpu_3d"):
        assert os.path.exists(os.path.join(ROOT, "bin", "build", name))
    if _has_gpu():
        return
    for name in ("test_nufi_gpu_3d", "test_nufi_cpu_2d"):
        r = subprocess.run([os.path.join(ROOT, "bin", "build", name), "--steps", "1"], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0
        assert "error:" in r.stderr and ("CUDA" in r.stderr or "cuda" in r.stderr)


REF_BIN = "/root/reference/bin"
REF_DRIVERS = ["test_nufi_cpu_1d", "test_nufi_cpu_2d", "test_nufi_cpu_3d", "test_nufi_gpu_2d", "test_nufi_gpu_3d"]


@pytest.mark.skipif(not os.path.isdir(REF_BIN), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("name", REF_DRIVERS)
def test_unmodified_reference_drivers_compile_against_include(name):
    """INTEGRATION.md section A, literally: the reference's OWN driver sources, unmodified and read where they lie, compile and
    link with -I<repo>/include ALONE (no reference include path behind it, no nvcc, FFTW, BLAS or MPI) against libnufi_b200.so.
    (bin/test_nufi_gpu_1d.cpp needs Armadillo for its SVD post-processing -- out of scope -- and is not in the list.)
    __graft_entry__.build() leaves the same binaries in oracle/_ref/drivers/ref_* so the GPU suite can run them."""
    out = os.path.join(ROOT, "oracle", "_ref", "drivers", "ref_" + name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-fopenmp", "-I" + os.path.join(ROOT, "include"), os.path.join(REF_BIN, name + ".cpp"),
                        "-o", out, "-L" + os.path.dirname(_lib.LIB_PATH), "-lnufi_b200", "-Wl,-rpath,$ORIGIN/../../../numericalflowiteration_b200/lib",
                        "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    if not _has_gpu():  # no silent CPU path behind the reference's own loop either
        rr = subprocess.run([out], capture_output=True, text=True, timeout=120, cwd=os.path.dirname(out))
        assert rr.returncode != 0 and ("CUDA" in rr.stderr or "cuda" in rr.stderr), rr.stderr[-500:]


def test_host_eval_any_order_and_derivative(tmp_path, oracle):
    """include/nufi/fields.hpp eval<real,order,dx[,dy,dz]> -- the host-side spline evaluation the drivers use for plots, generic
    in order and derivative like the reference's (nufi/fields.hpp:36-61; bin/test_nufi_gpu_1d.cpp:301 uses dx = 2) -- against
    the oracle's restatement of splines.hpp / fields.hpp (itself pinned bit for bit by tests/golden/orders.npz)."""
    src = r'''
#include <nufi/fields.hpp>
#include <cstdio>
#include <vector>
template <size_t O> void run(const std::vector<double>& l1, const std::vector<double>& l2) {
  nufi::dim1::config_t<double> a; a.Nx = 12; a.x_min = -1.0; a.x_max = 5.0; a.derive();
  nufi::dim2::config_t<double> b; b.Nx = 9; b.Ny = 8; b.y_min = 0.5; b.y_max = 7.0; b.derive();
  const double xs[3] = {-3.3, 0.2, 4.9999}, ys[3] = {0.1, 3.0, 9.7};
  for (int i = 0; i < 3; ++i) {
    std::printf("%.17g %.17g %.17g ", nufi::dim1::eval<double,O,0>(xs[i], l1.data(), a), nufi::dim1::eval<double,O,1>(xs[i], l1.data(), a),
                nufi::dim1::eval<double,O,2>(xs[i], l1.data(), a));
    std::printf("%.17g %.17g %.17g %.17g\n", nufi::dim2::eval<double,O,0,0>(xs[i], ys[i], l2.data(), b), nufi::dim2::eval<double,O,1,0>(xs[i], ys[i], l2.data(), b),
                nufi::dim2::eval<double,O,0,1>(xs[i], ys[i], l2.data(), b), nufi::dim2::eval<double,O,1,1>(xs[i], ys[i], l2.data(), b));
  }
}
int main() {
  std::vector<double> l1(12 + 7), l2((9 + 7) * (8 + 7));
  for (size_t i = 0; i < l1.size(); ++i) l1[i] = 0.3 + 0.01 * double((i * 7919) % 101);
  for (size_t i = 0; i < l2.size(); ++i) l2[i] = -0.2 + 0.01 * double((i * 104729) % 97);
  run<3>(l1, l2); run<4>(l1, l2); run<5>(l1, l2); run<6>(l1, l2); run<8>(l1, l2);
  return 0; }'''
    (tmp_path / "t.cpp").write_text(src)
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "t"), str(tmp_path / "t.cpp"),
                    "-L", os.path.dirname(_lib.LIB_PATH), "-lnufi_b200", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)], check=True)
    rows = [[float(v) for v in ln.split()] for ln in subprocess.run([str(tmp_path / "t")], capture_output=True, text=True, check=True).stdout.splitlines()]
    a = Config1D(Nx=12, x_min=-1.0, x_max=5.0)
    b = Config2D(Nx=9, Ny=8, y_min=0.5, y_max=7.0)
    l1 = np.array([0.3 + 0.01 * ((i * 7919) % 101) for i in range(19)])
    l2 = np.array([-0.2 + 0.01 * ((i * 104729) % 97) for i in range(16 * 15)])
    xs, ys = [-3.3, 0.2, 4.9999], [0.1, 3.0, 9.7]
    k = 0
    for order in (3, 4, 5, 6, 8):
        for i in range(3):
            want = [oracle.field(a, l1, (xs[i],), (d,), order=order) for d in (0, 1, 2)]
            want += [oracle.field(b, l2, (xs[i], ys[i]), d, order=order) for d in ((0, 0), (1, 0), (0, 1), (1, 1))]
            assert np.allclose(rows[k], want, rtol=1e-12, atol=1e-13), (order, i, rows[k], want)
            k += 1


def test_reference_layout_header_compiles_as_cxx():
    """config_t<double> of include/nufi/config.hpp is layout-identical to the C structs (static_assert in
    cuda_scheduler.hpp) and keeps the reference's defaults."""
    src = r'''
#include <nufi/cuda_scheduler.hpp>
#include <cstdio>
int main(){
  nufi::dim1::config_t<double> a; nufi::dim2::config_t<double> b; nufi::dim3::config_t<double> c;
  std::printf("%zu %zu %zu %zu %zu %zu %.17g %.17g %.17g\n", a.Nx, a.Nu, a.Nt, b.Nt, c.Nt, sizeof(c), a.dx, b.dv, c.x_max);
  return 0; }'''
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(src)
        subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.cpp"),
                        "-L", os.path.dirname(_lib.LIB_PATH), "-lnufi_b200", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    c1, c2, c3 = Config1D(), Config2D(), Config3D()
    assert [int(x) for x in out[:6]] == [256, 512, 1600, 800, 50, 35 * 8]
    assert float(out[6]) == c1.dx and float(out[7]) == c2.dv and float(out[8]) == c3.x_max


def test_history_file_formats_roundtrip(tmp_path):
    """Binary checkpoint and the reference's two text formats (isolated-step harness, 1d driver dump), Python and C++."""
    from numericalflowiteration_b200 import history_io

    conf = Config3D(Nx=5, Ny=4, Nz=6, Nt=3)
    st = stride_t(conf)
    rng = np.random.default_rng(5)
    coeffs = rng.standard_normal(4 * st) * 10.0 ** rng.integers(-12, 3, size=4 * st)
    history_io.write_binary(tmp_path / "h.bin", conf, coeffs, 4)
    hdr, back = history_io.read_binary(tmp_path / "h.bin")
    assert np.array_equal(back, coeffs) and hdr["n_levels"] == 4 and (hdr["Nx"], hdr["Ny"], hdr["Nz"]) == (5, 4, 6) and hdr["dt"] == conf.dt
    history_io.write_text_isolated(tmp_path / "h.txt", conf, coeffs, 4)
    lines = open(tmp_path / "h.txt").read().splitlines()
    assert lines[:7] == ["Nt = 3", f"dt = {conf.dt:f}", "Nx = 5", "Ny = 4", "Nz = 6", "order = 4", ""]  # isolated.cpp:64-70
    assert np.array_equal(history_io.read_text(tmp_path / "h.txt", 4 * st, header_lines=7), coeffs)
    history_io.write_text_plain(tmp_path / "p.txt", coeffs)
    assert np.array_equal(history_io.read_text(tmp_path / "p.txt", 4 * st), coeffs)
    # the C++ header reads what Python wrote and writes what Python reads
    src = r'''
#include <nufi/history_io.hpp>
#include <cstdio>
int main(int argc, char** argv){
  std::vector<double> c, t;
  auto h = nufi::history_io::read_binary(argv[1], c);
  nufi::history_io::read_text(argv[2], 7, c.size(), t);
  if (t != c) return 3;
  nufi::history_io::write_binary(argv[3], h, c.data());
  nufi::history_io::write_text_isolated(argv[4], h, c.data());
  std::printf("%u %u %llu %zu\n", h.dim, h.order, (unsigned long long)h.n_levels, nufi::history_io::stride_t(h));
  return 0; }'''
    open(tmp_path / "t.cpp", "w").write(src)
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "t"), str(tmp_path / "t.cpp")], check=True)
    out = subprocess.run([str(tmp_path / "t"), str(tmp_path / "h.bin"), str(tmp_path / "h.txt"), str(tmp_path / "h2.bin"), str(tmp_path / "h2.txt")],
                         capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [3, 4, 4, st]
    assert open(tmp_path / "h2.bin", "rb").read() == open(tmp_path / "h.bin", "rb").read()
    assert np.array_equal(history_io.read_text(tmp_path / "h2.txt", 4 * st, header_lines=7), coeffs)


def test_history_reader_does_not_trust_the_header(tmp_path):
    """A corrupt or foreign header (order 0, absurd sizes, a length the file does not have) is refused by both readers before it
    sizes an allocation; matches() compares a header with the running configuration."""
    import struct

    from numericalflowiteration_b200 import history_io

    conf = Config3D(Nx=5, Ny=4, Nz=6, Nt=3)
    st = stride_t(conf)
    good = tmp_path / "good.bin"
    history_io.write_binary(good, conf, np.arange(2 * st, dtype=np.float64), 2)
    raw = open(good, "rb").read()
    hdr = struct.Struct("<8sIIIIQQQQd")
    fields = list(hdr.unpack(raw[:64]))
    cases = {"order0": (3, 0), "dim7": (2, 7), "hugeNx": (5, 1 << 40), "levels": (8, 1 << 19), "truncated": None}
    src = r'''
#include <nufi/history_io.hpp>
int main(int argc, char** argv){
  std::vector<double> c;
  try { auto h = nufi::history_io::read_binary(argv[1], c); return nufi::history_io::matches(h, 3, 4, 5, 4, 6) ? 0 : 4; }
  catch (const std::runtime_error&) { return 7; } }'''
    open(tmp_path / "r.cpp", "w").write(src)
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "r"), str(tmp_path / "r.cpp")], check=True)
    assert subprocess.run([str(tmp_path / "r"), str(good)]).returncode == 0
    for name, edit in cases.items():
        f2 = list(fields)
        body = raw[64:]
        if edit is None:
            body = body[:-8]
        else:
            f2[edit[0]] = edit[1]
        path = tmp_path / f"{name}.bin"
        open(path, "wb").write(hdr.pack(*f2) + body)
        with pytest.raises(ValueError):
            history_io.read_binary(path)
        assert subprocess.run([str(tmp_path / "r"), str(path)]).returncode == 7, name


def test_c_abi_header_is_plain_c(tmp_path):
    """include/nufi_b200.h is the drop-in boundary: plain C (C99, -pedantic clean), no C++ or torch types in any signature."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "nufi_b200.h"\nint main(void) { nufi_b200_config3d c; (void)c; return NUFI_B200_PEER_HANDLE_BYTES == 64 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                        "-o", str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "nufi_b200.h")).read(), flags=re.S)  # declarations only
    assert "torch" not in hdr and "std::" not in hdr and "at::" not in hdr and "&" not in hdr


def test_fullsize_fixtures_are_consistent():
    """The committed full-size reference traces (tests/golden/fullsize_*.npz) belong to bench.py's workloads, start at the closed-form
    step-0 energy and, where both of the reference's builds were run, agree with each other while the problem is well conditioned."""
    import numpy as np

    sys.path.insert(0, ROOT)
    from bench import make_workload

    gold = os.path.join(ROOT, "tests", "golden")
    for name in ("C1", "C2", "C3", "C4"):
        g = np.load(os.path.join(gold, f"fullsize_{name}.npz"))
        conf, f0, _, desc = make_workload(name, 1)
        assert str(g["workload"]) == desc and int(g["f0_kind"]) == f0.kind and list(g["f0_p"]) == list(f0.p)
        assert len(g["energy"]) == int(g["steps"]) and np.all(np.isfinite(g["energy"])) and np.all(g["energy"] > 0)
    for name in ("C1", "C2"):  # 1d, alpha = 0.01, k = 0.5, L = 4 pi: E_0 = L alpha^2 / (4 k^2) = 4 pi * 1e-4
        a = np.load(os.path.join(gold, f"fullsize_{name}.npz"))["energy"]
        b = np.load(os.path.join(gold, f"fullsize_{name}_canonical.npz"))["energy"]
        assert abs(a[0] - 4 * np.pi * 1e-4) <= 1e-12 * a[0] and abs(b[0] - a[0]) <= 1e-12 * a[0]
        assert np.max(np.abs(a[:300] - b[:300]) / a[:300]) <= 1e-9  # -O3/FMA vs -O2 builds of the reference, first 300 steps


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the reference's own CPU eval_rho on the host cores) prints ONE JSON line with the keys the
    driver reads; it runs without a GPU."""
    import json

    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")  # what torch.distributed.run exports to rank 0
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C4", "--steps", "1",
                        "--warmup", "3", "--gpus", "2"], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "backtrace point-steps/sec" and d["unit"] == "point-steps/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64" and "workload" in d["config"]
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))  # all host cores although the launcher exported OMP_NUM_THREADS=1
    assert "free run of the reference CPU loop" in d["config"]["history"]  # the same problem the GPU arm free-runs, not a synthetic field
    # the other ranks of a torchrun launch exit 0 without work and without output
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C4", "--steps", "1", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, env=dict(env, RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_radix4_stockham_index_scheme():
    """numpy restatement of the index arithmetic of csrc/tail.cu: fft_pass (one radix-2 stage when log2(Nd) is odd, then radix-4
    Stockham stages, twiddle index r*k*Nd/(4 Ns), output index ((j >> ls) << (ls+2)) + k + r*Ns) against numpy's FFT -- the scheme
    was validated this way on the CPU before the kernel ran on the GPU (where tests/test_gpu_parity.py::test_field_tail checks it)."""
    import numpy as np

    def fft_pass(x, sign):
        nd = len(x)
        lg = nd.bit_length() - 1
        tw = np.exp(2j * np.pi * np.arange(nd) / nd)

        def ctw(u, w):
            return u * (np.conj(w) if sign < 0 else w)

        cur, oth = x.astype(complex).copy(), np.zeros(nd, complex)
        ns, ls = 1, 0
        if lg & 1:
            half = nd >> 1
            for j in range(half):
                oth[2 * j], oth[2 * j + 1] = cur[j] + cur[j + half], cur[j] - cur[j + half]
            cur, oth, ns, ls = oth, cur, 2, 1
        quarter, lq = nd >> 2, lg - 2
        while ns < nd:
            tshift = lq - ls
            for j in range(quarter):
                k = j & (ns - 1)
                m = k << tshift
                v0, v1, v2, v3 = cur[j], ctw(cur[j + quarter], tw[m]), ctw(cur[j + 2 * quarter], tw[2 * m]), ctw(cur[j + 3 * quarter], tw[3 * m])
                a, b, c, d = v0 + v2, v0 - v2, v1 + v3, v1 - v3
                idd = -1j * d if sign < 0 else 1j * d
                o = ((j >> ls) << (ls + 2)) + k
                oth[o], oth[o + ns], oth[o + 2 * ns], oth[o + 3 * ns] = a + c, b + idd, a - c, b - idd
            cur, oth, ns, ls = oth, cur, ns << 2, ls + 2
        return cur

    rng = np.random.default_rng(1)
    for nd in (8, 16, 32, 64, 128, 256):
        x = rng.normal(size=nd) + 1j * rng.normal(size=nd)
        assert np.max(np.abs(fft_pass(x, -1) - np.fft.fft(x))) <= 1e-12
        assert np.max(np.abs(fft_pass(x, +1) - np.fft.ifft(x) * nd)) <= 1e-12
