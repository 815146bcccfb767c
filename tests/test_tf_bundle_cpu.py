"""The TensorFlow tensor-bundle reader (tf_bundle.py) and the policy SavedModel loader (checkpoint.py):
SURVEY.md section 8f-3.  No TensorFlow here, so the reader is checked against published constants, hand-assembled
blocks and this repo's own writer (parity with a file written by the reference is UNPINNED, see tf_bundle.py)."""
import os
import struct

import numpy as np
import pytest

import cmbpo_b200 as cb
from cmbpo_b200 import tf_bundle as tb
from cmbpo_b200 import checkpoint as ck


def test_crc32c_published_check_values():
    assert tb.crc32c(b"123456789") == 0xE3069283                 # the CRC-32C "check" value (RFC 3720 appendix B.4 family)
    assert tb.crc32c(b"\0" * 32) == 0x8A9136AA                     # RFC 3720 B.4: 32 bytes of zeros
    assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43                   # RFC 3720 B.4: 32 bytes of ones
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E               # RFC 3720 B.4: incrementing bytes
    # masking: rotate right by 15, add the constant (lib/hash/crc32c.h)
    c = 0x12345678
    assert tb.mask_crc(c) == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_snappy_hand_assembled_stream():
    # "abcdabcdabcdabXY": literal "abcd", copy(offset 4, length 10) with a 1-byte offset tag, literal "XY"
    stream = bytes([16]) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((10 - 4) << 2) | 1, 4]) + bytes([(2 - 1) << 2]) + b"XY"
    assert tb.snappy_decompress(stream) == b"abcdabcdabcdabXY"
    # 2-byte-offset copy and a long literal (length byte follows the tag)
    lit = bytes(range(70))
    stream = tb._put_varint(70 + 5) + bytes([60 << 2, 69]) + lit + bytes([((5 - 1) << 2) | 2, 70, 0])
    assert tb.snappy_decompress(stream) == lit + lit[:5]
    with pytest.raises(ValueError):
        tb.snappy_decompress(bytes([4, ((4 - 1) << 2) | 2, 9, 0]))       # copy before any output


def test_block_prefix_compression_and_restarts():
    # hand-assembled block: "apple"->1, "apply"->2 (shares 4), restart, "banana"->3
    e1 = bytes([0, 5, 1]) + b"apple" + b"1"
    e2 = bytes([4, 1, 1]) + b"y" + b"2"
    e3 = bytes([0, 6, 1]) + b"banana" + b"3"
    body = e1 + e2 + e3 + struct.pack("<III", 0, len(e1) + len(e2), 2)
    assert tb._block_entries(body) == [(b"apple", b"1"), (b"apply", b"2"), (b"banana", b"3")]


@pytest.mark.parametrize("compress", [False, True])
def test_bundle_round_trip_many_blocks(tmp_path, compress):
    rng = np.random.default_rng(0)
    tens = {"AC/pi/dense_%d/kernel" % i: rng.standard_normal((7, 5)).astype(np.float32) for i in range(40)}
    tens["AC/counter"] = np.array(12345678901, np.int64)
    tens["AC/flags"] = np.array([True, False, True])
    tens["AC/zeros"] = np.zeros((3, 100), np.float64)             # long runs: the writer's back-references
    tens["AC/half"] = rng.standard_normal(9).astype(np.float16)
    prefix = str(tmp_path / "variables" / "variables")
    tb.write_bundle(prefix, tens, block_size=256, compress=compress)
    entries, header = tb.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and set(entries) == set(tens)
    got = tb.read_bundle(prefix)
    for k, v in tens.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    only = tb.read_bundle(prefix, names=["AC/flags"])
    assert list(only) == ["AC/flags"]
    with pytest.raises(KeyError):
        tb.read_bundle(prefix, names=["missing"])


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "v")
    tb.write_bundle(prefix, {"a": np.arange(10, dtype=np.float32)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        tb.read_bundle(prefix)
    assert tb.read_bundle(prefix, verify=False)["a"].shape == (10,)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        tb.read_index(prefix + ".index")


def _policy_variables(seed, O=17, A=6, E=3, H=128):
    from cmbpo_b200 import workload as wl
    dyn, actor, v, vc = wl.make_problem(seed, O, A, hidden=(64, 64), vf_hidden=(H, H), a_hidden=(H, H))
    tens = {}
    for i, (w, b) in enumerate(zip(actor.W, actor.b)):
        nm = "AC/pi/dense" + ("_%d" % i if i else "")
        tens[nm + "/kernel"] = w
        tens[nm + "/bias"] = np.asarray(b).reshape(-1)
        tens[nm + "/kernel/Adam"] = np.zeros_like(w)               # optimizer slots must be ignored
        tens[nm + "/kernel/Adam_1"] = np.zeros_like(w)
    tens["AC/pi/log_std"] = actor.log_std
    for name, ens in (("VEnsemble", v), ("VCEnsemble", vc), ("BNN", dyn)):
        scope = "AC/" + name if name != "BNN" else name
        for i, (w, b) in enumerate(zip(ens.W, ens.b)):
            tens["%s/Layer%d/FC_weights" % (scope, i)] = w
            tens["%s/Layer%d/FC_biases" % (scope, i)] = np.asarray(b).reshape(w.shape[0], 1, -1)
            tens["%s/Layer%d/FC_weights/Adam" % (scope, i)] = np.zeros_like(w)
        tens[scope + "/scaler_in_mu"] = ens.mu_in
        tens[scope + "/scaler_in_std"] = ens.var_in
        tens[scope + "/scaler_out_mu"] = ens.mu_out
        tens[scope + "/scaler_out_std"] = ens.var_out
        tens[scope + "/scaler_in_count"] = np.array(0.0, np.float32)
    tens["beta1_power"] = np.array(0.9, np.float32)
    return tens, actor, v, vc


def test_read_policy_savedmodel(tmp_path):
    tens, actor, v, vc = _policy_variables(3)
    export = str(tmp_path / "policy")
    tb.write_bundle(os.path.join(export, "variables", "variables"), tens, compress=True)
    got = ck.read_policy_savedmodel(export, vf_activation="swish", vf_elites=[0, 2])
    assert len(got["actor_W"]) == 3
    for a, b in zip(got["actor_W"], actor.W):
        assert np.array_equal(a, b)
    for a, b in zip(got["actor_b"], actor.b):
        assert np.array_equal(a, np.asarray(b).reshape(-1))
    assert np.array_equal(got["log_std"], actor.log_std)
    for key, ens in (("v", v), ("vc", vc)):
        e = got[key]
        assert e.acts == ["swish", "swish", None] and not e.probabilistic and e.elite_inds == [0, 2]
        for a, b in zip(e.W, ens.W):
            assert np.array_equal(a, b)
        for a, b in zip(e.b, ens.b):
            assert np.array_equal(a, np.asarray(b).reshape(a.shape))
        assert np.array_equal(e.var_out, ens.var_out) and np.array_equal(e.mu_in, ens.mu_in)
    # a SavedModel without an actor is refused
    tb.write_bundle(os.path.join(str(tmp_path / "empty"), "variables", "variables"), {"x": np.zeros(3, np.float32)})
    with pytest.raises(ValueError):
        ck.read_policy_savedmodel(str(tmp_path / "empty"))
