"""Test-side helpers (the product package never sees the oracle)."""
import numpy as np


def mirror_from_oracle(gpu, oracle_index):
    """Upload the rows of an oracle.hxo.Index (vectors, every graph layer, entry point) into a device VectorIndex."""
    ids = oracle_index.node_ids()
    rows = np.stack([oracle_index.vector(int(i)) for i in ids]) if len(ids) else np.zeros((0, gpu.dim), np.float32)
    gpu.load_vectors(ids, rows)
    graph, state = oracle_index.export_graph()
    for layer, (nodes, offs, nbrs) in graph.items():
        gpu.load_graph(layer, nodes, offs, nbrs)
    if state is not None:
        gpu.set_entry(state[0], state[1])


def build_and_run_cpp_selftest(tmp_path):
    """Compiles tests/cpp/host_mirror_selftest.cpp against helix-db_b200/host/vector_index.hpp, LINKS it with the in-tree
    libhelix_b200.so and runs it.  Returns (returncode, stdout + stderr)."""
    import os
    import subprocess
    import sysconfig
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    libdir = root / "helix-db_b200"
    exe = Path(tmp_path) / "host_mirror_selftest"
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", str(root / "tests" / "cpp" / "host_mirror_selftest.cpp"),
                         f"-I{libdir / 'host'}", f"-L{libdir}", "-lhelix_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)],
                        capture_output=True, text=True)
    if cc.returncode != 0:
        return cc.returncode, cc.stdout + cc.stderr
    env = dict(os.environ)
    extra = ["/usr/local/cuda/lib64", str(Path(sysconfig.get_paths()["purelib"]) / "nvidia" / "cuda_runtime" / "lib")]
    env["LD_LIBRARY_PATH"] = ":".join([env.get("LD_LIBRARY_PATH", "")] + extra).strip(":")
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=300)
    return run.returncode, run.stdout + run.stderr
