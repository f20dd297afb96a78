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
