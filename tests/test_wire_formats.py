"""CPU: the host-side wire formats of the C ABI (big-endian records, partial-witness records, Vec helpers, JSON, getters)
against the oracle's restatement of rln/src/utils.rs and rln/src/protocol/{witness,proof}.rs.  Mirrors the round-trip tests of
rln/tests/serialize.rs and the Shamir recovery of rln/tests/protocol.rs; nothing here launches a kernel."""
import json
import random

import pytest

from common import R, fr_stream
from pyref import keygen as K
from pyref import serialize as S


@pytest.fixture(scope="module")
def z():
    import zerokit_b200
    return zerokit_b200


def _witness_args(seed, depth=20):
    fs = fr_stream(seed)
    limit = 100
    return dict(secret=next(fs), limit=limit, mid=7, path=[next(fs) for _ in range(depth)], idx=[(3 * i + seed) % 2 for i in range(depth)],
                x=next(fs), en=next(fs))


def test_witness_be_le_and_getters(z):
    for seed, depth in ((1, 20), (2, 10), (3, 0)):
        a = _witness_args(seed, depth)
        w = z.RLNWitnessInput.new_single(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
        le = S.witness_to_bytes(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"], be=False)
        be = S.witness_to_bytes(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"], be=True)
        assert w.to_bytes_le() == le and w.to_bytes_be() == be
        assert z.RLNWitnessInput.from_bytes_be(be).to_bytes_le() == le
        assert z.RLNWitnessInput.from_bytes_le(le).to_bytes_be() == be
        assert (w.version_byte, w.identity_secret, w.user_message_limit, w.message_id) == (0, a["secret"], a["limit"], a["mid"])
        assert (w.path_elements, w.identity_path_index, w.x, w.external_nullifier) == (a["path"], a["idx"], a["x"], a["en"])
        assert w.to_bigint_json() == S.witness_to_bigint_json(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
        assert json.loads(w.to_bigint_json())["pathElements"] == [str(v) for v in a["path"]]
    # errors: truncated, trailing bytes, non-canonical element, unknown mode byte, message id out of range
    with pytest.raises(z.RLNError, match="too short"):
        z.RLNWitnessInput.from_bytes_be(be[:-5])
    with pytest.raises(z.RLNError, match="Expected to read"):
        z.RLNWitnessInput.from_bytes_be(be + b"\0")
    with pytest.raises(z.RLNError, match="Non-canonical field element"):
        z.RLNWitnessInput.from_bytes_be(be[:1] + R.to_bytes(32, "big") + be[33:])
    with pytest.raises(z.RLNError, match="Unknown message mode version byte: 0x07"):
        z.RLNWitnessInput.from_bytes_be(b"\x07" + be[1:])
    bad = S.witness_to_bytes(a["secret"], 5, 5, a["path"], a["idx"], a["x"], a["en"], be=True)
    with pytest.raises(z.RLNError, match="is not within user_message_limit"):
        z.RLNWitnessInput.from_bytes_be(bad)


def test_witness_multi_be(z):
    a = _witness_args(4, 20)
    mids, sel = [1, 2, 3, 0], [True, True, False, False]
    w = z.RLNWitnessInput.new_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel)
    le = S.witness_to_bytes_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel, be=False)
    be = S.witness_to_bytes_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel, be=True)
    assert w.to_bytes_le() == le and w.to_bytes_be() == be
    assert z.RLNWitnessInput.from_bytes_be(be).to_bytes_le() == le
    assert (w.version_byte, w.message_ids, w.selector_used) == (1, mids, sel)
    j = json.loads(w.to_bigint_json())
    assert j["messageId"] == [str(v) for v in mids] and j["selectorUsed"] == ["1", "1", "0", "0"]


def test_partial_witness_records(z):
    a = _witness_args(5, 20)
    pw = z.RLNPartialWitnessInput.new(a["secret"], a["limit"], a["path"], a["idx"])
    le = S.partial_witness_to_bytes(a["secret"], a["limit"], a["path"], a["idx"], be=False)
    be = S.partial_witness_to_bytes(a["secret"], a["limit"], a["path"], a["idx"], be=True)
    assert pw.to_bytes_le() == le and pw.to_bytes_be() == be
    assert z.RLNPartialWitnessInput.from_bytes_le(le).to_bytes_be() == be
    assert z.RLNPartialWitnessInput.from_bytes_be(be).to_bytes_le() == le
    assert (pw.version_byte, pw.identity_secret, pw.user_message_limit, pw.path_elements, pw.identity_path_index) == \
        (0, a["secret"], a["limit"], a["path"], a["idx"])
    # From<&RLNWitnessInput> (witness.rs:305-314)
    w = z.RLNWitnessInput.new_single(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
    assert w.to_partial().to_bytes_le() == le
    with pytest.raises(z.RLNError, match="Expected to read"):
        z.RLNPartialWitnessInput.from_bytes_le(le + b"\x01")
    with pytest.raises(z.RLNError, match="User message limit cannot be zero"):
        z.RLNPartialWitnessInput.from_bytes_le(S.partial_witness_to_bytes(a["secret"], 0, a["path"], a["idx"]))


def test_proof_values_be(z):
    fs = fr_stream(6)
    root, en, x, y, nul = (next(fs) for _ in range(5))
    le, be = S.proof_values_to_bytes(root, en, x, y, nul), S.proof_values_to_bytes(root, en, x, y, nul, be=True)
    assert z.proof_values_le_to_be(le) == be and z.proof_values_be_to_le(be) == le
    ys, nulls, sel = [next(fs) for _ in range(4)], [next(fs) for _ in range(4)], [1, 0, 1, 1]
    le = S.proof_values_to_bytes_multi(root, en, x, ys, nulls, sel)
    be = S.proof_values_to_bytes_multi(root, en, x, ys, nulls, sel, be=True)
    assert z.proof_values_le_to_be(le) == be and z.proof_values_be_to_le(be) == le
    with pytest.raises(z.RLNError, match="too short"):
        z.proof_values_be_to_le(be[:-1])


def test_vec_helpers(z):
    fs = fr_stream(7)
    for n in (0, 1, 5):
        vals = [next(fs) for _ in range(n)]
        for be in (False, True):
            b = z.vec_fr_to_bytes(vals, be)
            assert b == S.vec_fr(vals, be) and z.bytes_to_vec_fr(b, be) == vals
        raw = bytes(range(n))
        for be in (False, True):
            b = z.vec_u8_to_bytes(raw, be)
            assert b == S.vec_u8(raw, be) and z.bytes_to_vec_u8(b, be) == raw
    with pytest.raises(z.RLNError):
        z.bytes_to_vec_fr(S.vec_fr([1, 2], True)[:-1], True)
    with pytest.raises(z.RLNError, match="Non-canonical"):
        z.bytes_to_vec_fr(S.vec_fr([R], True), True)
    with pytest.raises(z.RLNError):
        z.bytes_to_vec_u8(b"\x09" + b"\0" * 7 + b"abc", False)


def test_compute_and_recover_id_secret(z):
    """rln/src/protocol/slashing.rs; property test as in rln/tests/protocol.rs (two shares of one line → a0)"""
    rnd = random.Random(8)
    for _ in range(20):
        a0, a1, x1, x2 = (rnd.randrange(R) for _ in range(4))
        s1, s2 = (x1, (a0 + x1 * a1) % R), (x2, (a0 + x2 * a1) % R)
        assert z.compute_id_secret(s1, s2) == a0 == K.compute_id_secret(s1, s2)
    with pytest.raises(z.RLNError, match="division by zero"):
        z.compute_id_secret((5, 1), (5, 2))
    a0, a1, en, root, nul = (rnd.randrange(R) for _ in range(5))
    pv = [S.proof_values_to_bytes(root, en, x, (a0 + x * a1) % R, nul) for x in (11, 12)]
    assert z.recover_id_secret(pv[0], pv[1]) == a0
    other = S.proof_values_to_bytes(root, (en + 1) % R, 13, (a0 + 13 * a1) % R, nul)
    with pytest.raises(z.RLNError, match="External nullifiers mismatch"):
        z.recover_id_secret(pv[0], other)
    # multi: the first pair of used slots sharing a nullifier gives the shares
    m1 = S.proof_values_to_bytes_multi(root, en, 21, [(a0 + 21 * a1) % R, 5], [nul, 77], [1, 1])
    m2 = S.proof_values_to_bytes_multi(root, en, 22, [9, (a0 + 22 * a1) % R], [78, nul], [1, 1])
    assert z.recover_id_secret(m1, m2) == a0
    m3 = S.proof_values_to_bytes_multi(root, en, 22, [9, (a0 + 22 * a1) % R], [78, nul], [1, 0])
    with pytest.raises(z.RLNError, match="No matching nullifier"):
        z.recover_id_secret(m1, m3)
    with pytest.raises(z.RLNError, match="No matching nullifier"):
        z.recover_id_secret(pv[0], m2)


# ------------------------------------------------------------------------------------------------ V3 records (ffi_rln_v3.rs)
def test_v3_witness_records(z):
    """rln/tests/serialize.rs V3 round trips; layouts from rln/src/protocol/serialize.rs (the BE Single order differs from LE)"""
    a = _witness_args(11, 20)
    w = z.WitnessV3.new_single(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
    le = S.v3_witness_single(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
    be = S.v3_witness_single(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"], be=True)
    assert w.to_bytes_le() == le and w.to_bytes_be() == be
    assert z.WitnessV3.from_bytes_le(le).to_bytes_be() == be and z.WitnessV3.from_bytes_be(be).to_bytes_le() == le
    assert (w.identity_secret, w.user_message_limit, w.message_id, w.message_ids, w.selector_used) == (a["secret"], a["limit"], a["mid"], None, None)
    assert (w.path_elements, w.identity_path_index, w.x, w.external_nullifier) == (a["path"], a["idx"], a["x"], a["en"])
    mids, sel = [1, 2, 3, 0], [True, False, True, False]
    m = z.WitnessV3.new_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel)
    le = S.v3_witness_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel)
    be = S.v3_witness_multi(a["secret"], a["limit"], mids, a["path"], a["idx"], a["x"], a["en"], sel, be=True)
    assert m.to_bytes_le() == le and m.to_bytes_be() == be
    assert z.WitnessV3.from_bytes_le(le).to_bytes_be() == be and z.WitnessV3.from_bytes_be(be).to_bytes_le() == le
    assert (m.message_id, m.message_ids, m.selector_used) == (None, mids, sel)
    # constructor errors carry the V3 texts (rln/src/error.rs:136-167)
    with pytest.raises(z.RLNError, match="Field `path_elements` has length 20, but field `identity_path_index` has length 19"):
        z.WitnessV3.new_single(1, 10, 1, a["path"], a["idx"][:-1], 1, 1)
    with pytest.raises(z.RLNError, match="At least one value in `selector_used` must be true"):
        z.WitnessV3.new_multi(1, 10, [1, 2], a["path"], a["idx"], 1, 1, [False, False])
    with pytest.raises(z.RLNError, match="Duplicate message ID found in `message_ids`"):
        z.WitnessV3.new_multi(1, 10, [2, 2], a["path"], a["idx"], 1, 1, [True, True])
    with pytest.raises(z.RLNError, match="Field `message_ids` has length 2, but field `selector_used` has length 1"):
        z.WitnessV3.new_multi(1, 10, [1, 2], a["path"], a["idx"], 1, 1, [True])
    with pytest.raises(z.RLNError, match="failed to fill whole buffer"):
        z.WitnessV3.from_bytes_le(le[:-1])
    with pytest.raises(z.RLNError, match="Non-canonical field element"):
        z.WitnessV3.from_bytes_be(be[:1] + R.to_bytes(32, "big") + be[33:])
    # partial witness: no tag byte
    pw = z.PartialWitnessV3.new(a["secret"], a["limit"], a["path"], a["idx"])
    ple, pbe = S.v3_partial_witness(a["secret"], a["limit"], a["path"], a["idx"]), S.v3_partial_witness(a["secret"], a["limit"], a["path"], a["idx"], be=True)
    assert pw.to_bytes_le() == ple and pw.to_bytes_be() == pbe and w.to_partial().to_bytes_le() == ple
    assert z.PartialWitnessV3.from_bytes_le(ple).to_bytes_be() == pbe and z.PartialWitnessV3.from_bytes_be(pbe).to_bytes_le() == ple
    assert (pw.identity_secret, pw.user_message_limit, pw.path_elements, pw.identity_path_index) == (a["secret"], a["limit"], a["path"], a["idx"])


def test_v3_proof_values_records(z):
    fs = fr_stream(12)
    y, root, nul, x, en = (next(fs) for _ in range(5))
    le, be = S.v3_values_single(y, root, nul, x, en), S.v3_values_single(y, root, nul, x, en, be=True)
    v = z.ProofValuesV3.from_bytes_le(le)
    assert v.to_bytes_le() == le and v.to_bytes_be() == be and z.ProofValuesV3.from_bytes_be(be).to_bytes_le() == le
    assert (v.y, v.root, v.nullifier, v.x, v.external_nullifier, v.ys, v.nullifiers, v.selector_used) == (y, root, nul, x, en, None, None, None)
    ys, nulls, sel = [next(fs) for _ in range(4)], [next(fs) for _ in range(4)], [True, True, False, True]
    le, be = S.v3_values_multi(ys, root, nulls, x, en, sel), S.v3_values_multi(ys, root, nulls, x, en, sel, be=True)
    m = z.ProofValuesV3.from_bytes_le(le)
    assert m.to_bytes_le() == le and m.to_bytes_be() == be and z.ProofValuesV3.from_bytes_be(be).to_bytes_le() == le
    assert (m.y, m.nullifier, m.ys, m.nullifiers, m.selector_used) == (None, None, ys, nulls, sel)
    with pytest.raises(z.RLNError, match="Non-canonical bool byte: expected 0x00 or 0x01, got 0x02"):
        z.ProofValuesV3.from_bytes_be(be[:-1] + b"\x02")
    with pytest.raises(z.RLNError, match="failed to fill whole buffer"):
        z.ProofValuesV3.from_bytes_be(be[:-1])
    # Shamir recovery through the V3 names
    a0, a1 = next(fs), next(fs)
    v1 = z.ProofValuesV3.from_bytes_le(S.v3_values_single((a0 + 5 * a1) % R, root, nul, 5, en))
    v2 = z.ProofValuesV3.from_bytes_le(S.v3_values_single((a0 + 6 * a1) % R, root, nul, 6, en))
    assert v1.recover_id_secret(v2) == a0 == z.compute_id_secret_v3((5, (a0 + 5 * a1) % R), (6, (a0 + 6 * a1) % R))


def test_codec_round_trips_random_shapes(z):
    """property test in the spirit of rln/tests/serialize.rs: random depths, values at the field boundary, every record type,
    LE ↔ BE ↔ V3 conversions all return to the same bytes"""
    rnd = random.Random(99)
    edge = [0, 1, R - 1, R - 2, 1 << 253, (1 << 128) - 1]
    for _ in range(40):
        depth = rnd.choice([0, 1, 2, 7, 20, 32])
        pick = lambda: rnd.choice(edge + [rnd.randrange(R)])
        limit = rnd.choice([1, 2, 100, R - 1])
        mid = rnd.randrange(limit)
        a = dict(secret=pick(), limit=limit, mid=mid, path=[pick() for _ in range(depth)], idx=[rnd.randrange(2) for _ in range(depth)],
                 x=pick(), en=pick())
        w = z.RLNWitnessInput.new_single(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
        le, be = w.to_bytes_le(), w.to_bytes_be()
        assert le == S.witness_to_bytes(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
        assert z.RLNWitnessInput.from_bytes_le(le).to_bytes_be() == be and z.RLNWitnessInput.from_bytes_be(be).to_bytes_le() == le
        v3 = z.WitnessV3.from_bytes_be(be)            # the V1 and V3 BE records are the same bytes
        assert v3.to_bytes_be() == be and z.WitnessV3.from_bytes_le(v3.to_bytes_le()).to_bytes_be() == be
        assert v3.to_bytes_le() == S.v3_witness_single(a["secret"], a["limit"], a["mid"], a["path"], a["idx"], a["x"], a["en"])
        pw = w.to_partial()
        assert z.RLNPartialWitnessInput.from_bytes_be(pw.to_bytes_be()).to_bytes_le() == pw.to_bytes_le()
        assert pw.to_bytes_le()[1:] == v3.to_partial().to_bytes_le()      # V3 drops the version byte
        k = rnd.choice([1, 2, 4, 8])
        ys, nulls, sel = [pick() for _ in range(k)], [pick() for _ in range(k)], [rnd.random() < 0.6 for _ in range(k)]
        root = pick()
        m_le = S.proof_values_to_bytes_multi(root, a["en"], a["x"], ys, nulls, sel)
        assert z.proof_values_be_to_le(z.proof_values_le_to_be(m_le)) == m_le
        v3v = z.ProofValuesV3.from_bytes_le(S.v3_values_multi(ys, root, nulls, a["x"], a["en"], sel))
        assert z.ProofValuesV3.from_bytes_be(v3v.to_bytes_be()).to_bytes_le() == v3v.to_bytes_le()
        assert (v3v.ys, v3v.nullifiers, v3v.selector_used, v3v.root) == (ys, nulls, sel, root)
