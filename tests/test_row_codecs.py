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
