"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/*.h declares;
host-side layout helpers; no compute calls (no GPU here)."""
import os
import re
import subprocess

import numpy as np
import pytest

import diffopt_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "diffopt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(diffopt_b200_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    capi = diffopt_b200.submodule("_capi")
    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/diffopt_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(capi.SIGNATURES) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (diffopt_b200_\w+)", out))
    assert exported == set(declared)


def test_library_carries_sm100a_code():
    capi = diffopt_b200.submodule("_capi")
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(diffopt_b200.DiffOptB200Error):
        diffopt_b200.Context(0)


def test_colmajor_roundtrip():
    qp = diffopt_b200.submodule("qp")
    X = np.arange(2 * 3 * 4, dtype=float).reshape(2, 3, 4)
    buf = qp.colmajor(X, 3, 4, 2)
    assert buf.flags.c_contiguous and buf.shape == (2, 4, 3)
    # column-major per instance: element (i, j) of instance b at b*12 + j*3 + i
    flat = buf.ravel()
    assert flat[1 * 12 + 2 * 3 + 1] == X[1, 1, 2]
    assert np.array_equal(qp.from_colmajor(buf, 3, 4), X)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "diffopt.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "").replace("the oracle", "") or f == "__init__.py", f


def _build_c99_client(tmpdir):
    capi = diffopt_b200.submodule("_capi")
    capi.load()
    exe = os.path.join(str(tmpdir), "c99_client")
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "c99_client.c"), "-o", exe, "-L", libdir, "-ldiffopt_b200",
           f"-Wl,-rpath,{libdir}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c99_and_a_c_client_links(tmp_path):
    """include/diffopt_b200.h compiled as C99 (-pedantic -Werror) in a translation unit that calls
    create -> qp_batch_solve -> destroy and links against the shared library.  On a box without a GPU the client runs up
    to the clean refusal of diffopt_b200_create (no CPU fallback); the GPU run of the same binary is in test_qp_gpu.py."""
    exe = _build_c99_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


def test_coo_batch_packing_matches_the_references_triplets():
    """qp.coo_batch: per-instance scipy.sparse matrices -> 0-based offsets + Julia's 1-based (I, J, V), duplicates kept as
    separate triplets (the device adds them up like SparseArrays.sparse, QuadraticProgram.jl:396-424)."""
    import scipy.sparse as sp
    qp = diffopt_b200.submodule("qp")
    m0 = sp.coo_matrix((np.array([1.5, -2.0]), (np.array([0, 2]), np.array([1, 0]))), shape=(3, 4))
    m1 = sp.coo_matrix((np.array([0.25, 0.75]), (np.array([1, 1]), np.array([3, 3]))), shape=(3, 4))   # a duplicated entry
    m2 = sp.csr_matrix((3, 4))
    ptr, I, J, V = qp.coo_batch([m0, m1, m2], 3)
    assert ptr.dtype == np.int64 and I.dtype == np.int64 and J.dtype == np.int64 and V.dtype == np.float64
    assert ptr.tolist() == [0, 2, 4, 4]
    assert sorted(zip(I[:2].tolist(), J[:2].tolist(), V[:2].tolist())) == [(1, 2, 1.5), (3, 1, -2.0)]
    assert I[2:].tolist() == [2, 2] and J[2:].tolist() == [4, 4] and V[2:].tolist() == [0.25, 0.75]
    ptr, I, J, V = qp.coo_batch(m0, 5, shared=True)
    assert ptr.tolist() == [0, 2]
    assert qp.coo_batch(None, 3) is None


def test_sparse_analysis_assembly_nodes_and_row_ordered_lists(monkeypatch):
    """Host-only analysis of the multifrontal path (`diffopt_b200_sparse_analyze`, no GPU): a top front with a thousand
    children gets two levels of assembly nodes (fronts without pivots: same factors, more fronts, two more launches);
    a chain-structured KKT pattern (config 3 at reduced horizon) is left alone."""
    import bench_data
    lsq = diffopt_b200.submodule("lsqr")
    K = bench_data.portfolio_config3(n=16000, nfac=20, density=0.2)["K"]
    with_nodes = lsq.sparse_analyze(K, trans=True)
    monkeypatch.setenv("DIFFOPT_B200_MF_NO_ASSEMBLY_NODES", "1")
    plain = lsq.sparse_analyze(K, trans=True)
    assert plain["levels"] == 2 and with_nodes["levels"] == 4
    assert with_nodes["fronts"] == plain["fronts"] + 42 + 2          # 1000 leaves / 24 -> 42 nodes / 24 -> 2 nodes
    assert with_nodes["nnz_lu"] == plain["nnz_lu"] and with_nodes["factor_flops"] == plain["factor_flops"]
    Km = bench_data.mpc_config3(T=300)["K"]
    a = lsq.sparse_analyze(Km, trans=True)
    monkeypatch.delenv("DIFFOPT_B200_MF_NO_ASSEMBLY_NODES")
    b = lsq.sparse_analyze(Km, trans=True)
    assert {k: v for k, v in a.items() if k != "analysis_ms"} == {k: v for k, v in b.items() if k != "analysis_ms"}
