"""The host restatement of the in-kernel noise stream (oracle/philox_ref.py) against the published Philox4x32-10
known-answer vectors (Random123 kat_vectors) and basic statistics of the normals."""
import numpy as np

import philox_ref


def _kat(ctr, key):
    return philox_ref.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0].tolist()


def test_philox4x32_10_known_answers():
    assert _kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _kat([f, f, f, f], [f, f]) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normal_stream_properties():
    x = philox_ref.philox_normal(1234, 200001)
    assert x.shape == (200001,) and x.dtype == np.float32 and np.isfinite(x).all()
    assert abs(float(x.mean())) < 0.01 and abs(float(x.std()) - 1.0) < 0.01
    assert np.array_equal(x[:1000], philox_ref.philox_normal(1234, 1000))        # a prefix is a prefix
    assert not np.array_equal(x[:1000], philox_ref.philox_normal(1235, 1000))    # the seed matters
