#!/usr/bin/env python
"""Byte-level fuzz of the product's row / key decoders (hx_decode_neighbor_row, hx_parse_vector_key, hx_encode_vector_key):
they parse bytes that come out of a KV store, so truncated, oversized and random inputs must be REJECTED (or decoded), never
read out of bounds.  Device-free; meant to run on the host-sanitized build of the library
(HELIX_B200_LIB=helix-db_b200/_variants/libhelix_b200_asan.so with libasan preloaded: scripts/product_host_sanitize.sh)."""
import ctypes as C
import struct
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import helix_db_b200 as hx  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    lib = hx.load_library()
    decoded = rejected = 0
    seeds = [bytes([0x12, 0, 0, 0, 2]) + struct.pack(">QQ", 3, 7),
             bytes([0x13, 0x01, 0, 0, 0, 2]) + struct.pack("<Q", 0x0102030405060708) + struct.pack(">QQ", 1, 9),
             bytes([0x13, 0, 0, 0, 0, 2]) + struct.pack(">QQ", 1, 9), struct.pack(">I", 2) + struct.pack(">QQ", 1, 2), b""]
    for it in range(40_000):
        base = bytearray(seeds[it % len(seeds)])
        mode = it % 7
        if mode == 0 and base:
            del base[int(rng.integers(0, len(base))):]                       # truncate
        elif mode == 1:
            base += bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))   # trailing garbage
        elif mode == 2 and base:
            base[int(rng.integers(0, len(base)))] = int(rng.integers(0, 256))    # one flipped byte
        elif mode == 3 and len(base) >= 5:
            base[1:5] = struct.pack(">I", int(rng.integers(0, 2**32)))           # absurd count field
        elif mode == 4:
            base = bytearray(rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8).tobytes())
        layer = int(rng.integers(0, 4)) if mode != 5 else 0
        try:
            ids, sim = hx.decode_neighbor_row(layer, bytes(base))
            decoded += 1
            # like the reference's decode_node_ids (values/vectors.rs:81-89) the decoder returns the ids AS STORED; only the
            # layer-0 encoder canonicalises (sorted, unique; the upper-layer codec stores the list as given, neighbors.rs:57-76),
            # and hx_index_load_neighbor_rows refuses a non-canonical row
            if ids:
                again = hx.decode_neighbor_row(layer, hx.encode_neighbor_row(layer, ids))[0]
                assert again == (sorted(set(ids)) if layer == 0 else ids)
        except hx.HelixDbError:
            rejected += 1
    keys_ok = keys_bad = 0
    valid = [hx.encode_vector_key(hx.VectorKey(kind, 0x1122334455667788, 42, 3 if kind == hx.KeyKind.UpperNeighbors else 0,
                                              0xABCD if kind == hx.KeyKind.Vector else 0))
             for kind in hx.KeyKind if kind != hx.KeyKind.Other]
    for it in range(40_000):
        raw = bytearray(valid[it % len(valid)]) if it % 2 else bytearray(rng.integers(0, 256, int(rng.integers(0, 48)), dtype=np.uint8).tobytes())
        mode = it % 5
        if mode == 0 and raw:
            del raw[int(rng.integers(0, len(raw))):]
        elif mode == 1 and raw:
            raw[int(rng.integers(0, len(raw)))] = int(rng.integers(0, 256))
        elif mode == 2:
            raw += bytes(rng.integers(0, 256, int(rng.integers(1, 12)), dtype=np.uint8))
        try:
            k = hx.parse_vector_key(bytes(raw))
            keys_ok += 1
            if k.kind != hx.KeyKind.Other:                                        # a parsed key re-encodes to the same bytes
                assert hx.encode_vector_key(k) == bytes(raw), (k, bytes(raw).hex())
        except hx.HelixDbError:
            keys_bad += 1
    print(f"neighbour rows: {decoded} decoded, {rejected} rejected; keys: {keys_ok} parsed, {keys_bad} rejected; no sanitizer report")


if __name__ == "__main__":
    main()
