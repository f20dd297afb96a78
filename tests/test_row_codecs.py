"""Row-image codecs against the reference's frozen byte layouts. No GPU needed."""
import struct

import pytest

import helix_db_b200 as hx


def be64(x):
    return struct.pack(">Q", x)


def test_upper_neighbor_bytes_are_frozen():
    # encoding/v1/values/vectors/neighbors.rs:130-139
    expected = struct.pack(">I", 2) + be64(1) + be64(2)
    assert hx.encode_neighbor_row(1, [1, 2]) == expected
    assert hx.decode_neighbor_row(1, expected) == ([1, 2], None)
    umax = (1 << 64) - 1
    assert hx.decode_neighbor_row(3, hx.encode_neighbor_row(3, [1, umax]))[0] == [1, umax]


def test_layer0_neighbor_bytes_are_frozen_and_canonical():
    # encoding/v1/values/vectors.rs:218-240
    assert hx.encode_neighbor_row(0, []) == bytes([0x12, 0, 0, 0, 0])
    assert hx.decode_neighbor_row(0, bytes([0x12, 0, 0, 0, 0])) == ([], None)
    expected = bytes([0x12, 0, 0, 0, 2]) + be64(3) + be64(7)
    assert hx.encode_neighbor_row(0, [7, 3, 7]) == expected            # sorted + deduplicated
    assert hx.decode_neighbor_row(0, expected) == ([3, 7], None)
    assert hx.decode_neighbor_row(0, b"") == ([], None)                # missing row == empty row


def test_layer0_record_with_and_without_simhash():
    # encoding/v1/values/vectors.rs:242-275
    rec = bytes([0x13, 0x01, 0, 0, 0, 2]) + struct.pack("<Q", 0x0102030405060708) + be64(1) + be64(9)
    assert hx.decode_neighbor_row(0, rec) == ([1, 9], 0x0102030405060708)
    rec2 = bytes([0x13, 0, 0, 0, 0, 2]) + be64(1) + be64(9)
    assert hx.decode_neighbor_row(0, rec2) == ([1, 9], None)


@pytest.mark.parametrize("layer,row", [
    (0, bytes([0x12, 0, 0, 0, 2]) + be64(3)),                   # count says 2, one id present
    (0, bytes([0x12, 0, 0, 0, 1]) + be64(3) + b"\x00"),          # trailing byte
    (0, bytes([0x07, 0, 0, 0, 0])),                              # unknown encoding type
    (0, bytes([0x13, 0x02, 0, 0, 0, 0])),                        # invalid flags
    (1, struct.pack(">I", 1)),                                   # upper row shorter than its count
    (1, b"\x00\x00"),                                            # shorter than the count prefix
])
def test_corrupt_rows_are_rejected(layer, row):
    with pytest.raises(hx.HelixDbError) as e:
        hx.decode_neighbor_row(layer, row)
    assert e.value.variant == "InvariantViolation"


# ---- row keys: the exact layouts pinned by encoding/v1/keys/vectors.rs tests (:1798-2000) --------------------------------------
INDEX_ID, ORDER_CODE, NODE_ID, LAYER = 0x0102030405060708, 0x1112131415161718, 0x2122232425262728, 0x4243


def _be(v, n=8):
    return v.to_bytes(n, "big")


def test_vector_keys_have_the_reference_layouts():
    K, KK = hx.VectorKey, hx.KeyKind
    cases = [
        (K(KK.Vector, INDEX_ID, NODE_ID, order_code=ORDER_CODE), b"\xF1" + _be(INDEX_ID) + b"\x02" + _be(ORDER_CODE) + _be(NODE_ID)),
        (K(KK.Layer0Neighbors, INDEX_ID, NODE_ID), b"\xF0" + _be(INDEX_ID) + b"\x16" + _be(NODE_ID)),
        (K(KK.UpperNeighbors, INDEX_ID, NODE_ID, layer=LAYER), b"\xF0" + _be(INDEX_ID) + b"\x11" + _be(LAYER, 2) + _be(NODE_ID)),
        (K(KK.SimHash, INDEX_ID, NODE_ID), b"\xF0" + _be(INDEX_ID) + b"\x12" + _be(NODE_ID)),
        (K(KK.UpperVector, INDEX_ID, NODE_ID), b"\xF0" + _be(INDEX_ID) + b"\x13" + _be(NODE_ID)),
        (K(KK.Metadata, INDEX_ID), b"\x03\x03" + _be(INDEX_ID) + b"\x01"),
    ]
    for key, raw in cases:
        assert hx.encode_vector_key(key) == raw
        assert hx.parse_vector_key(raw) == key


def test_vector_key_parser_skips_other_families_and_rejects_corruption():
    KK = hx.KeyKind
    others = [
        b"\xF0" + _be(INDEX_ID),                                             # hot-lane prefix
        b"\xF1" + _be(INDEX_ID),                                             # l0-lane prefix
        b"\xF1" + _be(INDEX_ID) + b"\x02",                                   # vector item prefix
        b"\xF1" + _be(INDEX_ID) + b"\x17" + _be(ORDER_CODE) + _be(NODE_ID),  # simhash directory
        b"\xF1" + _be(INDEX_ID) + b"\x04" + _be(0xFFFF - LAYER, 2) + _be(NODE_ID),  # entry candidate (sorted)
        b"\xF1" + _be(INDEX_ID) + b"\x05" + _be(NODE_ID),                    # entry candidate (node)
        b"\xF1" + _be(INDEX_ID) + b"\x15" + _be(NODE_ID) + _be(LAYER, 2) + _be(NODE_ID),  # reverse edge
        b"\x03\x03" + _be(INDEX_ID) + b"\x09",                               # txn guard
        b"\x03\x03" + _be(INDEX_ID),                                         # index prefix
    ]
    for raw in others:
        k = hx.parse_vector_key(raw)
        assert k.kind == KK.Other and k.index_id == INDEX_ID
    bad = [
        b"", b"\xAA", b"\xF0" + _be(INDEX_ID)[:4],
        b"\xF0" + _be(INDEX_ID) + b"\x16" + _be(NODE_ID) + b"\x00",          # trailing byte
        b"\xF0" + _be(INDEX_ID) + b"\x11" + _be(NODE_ID),                    # upper row without its layer
        b"\xF0" + _be(INDEX_ID) + b"\x77" + _be(NODE_ID),                    # unknown kind
        b"\xF1" + _be(INDEX_ID) + b"\x02" + _be(NODE_ID),                    # vector key without order code
        b"\x03\x04" + _be(INDEX_ID) + b"\x01",                               # not the vector index type
        b"\x03\x03" + _be(INDEX_ID) + b"\x02",                               # invalid default kind
    ]
    for raw in bad:
        with pytest.raises(hx.HelixDbError):
            hx.parse_vector_key(raw)
