"""The oracle (oracle/oracle_core.cpp) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py): dot sets, packed multiplicons of every list pair and
of every profile x list search -- bit-exact including p-values (same libm)."""
import numpy as np
import pytest

import parity_lib as pl
from iadhore_b200 import capi
from trace_io import read_trace


@pytest.mark.parametrize("path", pl.golden_traces(), ids=lambda p: p.split("/")[-1].split(".")[0])
def test_oracle_replays_reference_trace(path):
    tr = read_trace(path)
    eng = pl.oracle_engine(tr.dataset)
    errs = pl.replay_trace(eng, tr, rtol_p=0.0)
    assert not errs, errs[:5]


def test_golden_covers_the_quirks():
    """The golden set must exercise twisted clusters, false anchor points, both orientations
    and all three multiple-hypothesis modes, or parity on it proves little."""
    seen = dict(twisted=0, false_ap=0, minus=0, plus=0, hl_levels=set(), modes=set(), multi_cluster=0)
    for path in pl.golden_traces():
        tr = read_trace(path)
        seen["modes"].add(tr.params["mult_hyp"])
        for ev in tr.events:
            if ev[0] in ("l2", "hl"):
                for m in capi.parse_packed(ev[1].packed):
                    seen["hl_levels"].add(m.level)
                    seen["multi_cluster"] += len(m.clusters) > 1
                    for c in m.clusters:
                        seen["twisted"] += c.was_twisted
                        seen["false_ap"] += int((c.aps[:, 4] == 0).sum())
                        seen["plus" if c.orient else "minus"] += 1
    assert seen["modes"] == {0, 1, 2}
    assert seen["twisted"] > 0 and seen["false_ap"] > 0 and seen["minus"] > 0 and seen["plus"] > 0
    assert seen["multi_cluster"] > 0 and max(seen["hl_levels"]) >= 4


def test_packed_layout_sizes():
    """Wire sizes of the reference: AP 17 B, cluster 102 B + 17 n, multiplicon 33 B."""
    tr = read_trace(pl.golden_traces()[0])
    for ev in tr.events:
        if ev[0] == "l2" and len(ev[1].packed) > 4:
            ms = capi.parse_packed(ev[1].packed)
            want = 4 + sum(33 + sum(102 + 17 * len(c.aps) for c in m.clusters) for m in ms)
            assert want == len(ev[1].packed)
            return
    pytest.fail("no multiplicon in golden trace")


def test_gap_schedule():
    # Settings::getGapSizes for the shipped gap_size=30
    assert capi.gap_sizes(30) == [3, 4, 5, 6, 8, 11, 14, 18, 23, 30]
    tr = read_trace(pl.golden_traces()[0])
    assert capi.gap_sizes(tr.params["gap_size"]) == tr.params["gaps"]
    for g in (3, 5, 9, 10, 17, 35, 100):
        import ctypes as C
        lib = C.CDLL(capi.ORACLE_LIB)
        buf = (C.c_int32 * 10)()
        n = lib.oracle_gap_sizes(C.c_int(g), buf)
        assert capi.gap_sizes(g) == list(buf)[:n]


def test_omxn_known_answers():
    """hpmath::omxn KATs harvested from the reference objects."""
    tr = read_trace(pl.golden_traces()[0])
    eng = pl.oracle_engine(tr.dataset)
    out = eng.omxn(np.array([1e-4, 0.3]), np.array([1800, 7], np.int32))
    assert out[0].hex() == "0x1.aba78d4067e56p-1"
    assert out[1].hex() == "0x1.5152be12f5a5ep-4"


def test_stream_digest_separates_p_values_from_the_rest():
    """parity_lib.stream_digest / digest_mismatch (the known-answer form of the parity bar used at full bench size):
    a last-bit change of a p-value is tolerated, any other byte is not."""
    import struct
    tr = read_trace(pl.golden_traces()[0])
    streams = [ev[1].packed for ev in tr.events if ev[0] == "l2" and len(ev[1].packed) > 4][:5]
    want = pl.stream_digest(streams)
    assert want["multiplicons"] > 0 and want["clusters"] >= want["multiplicons"]
    assert pl.digest_mismatch(pl.stream_digest(streams), want) is None
    b = bytearray(streams[0])
    o = 4 + 33                                   # first cluster of the first multiplicon: p-value at offset 0
    (pv,) = struct.unpack_from("<d", b, o)
    struct.pack_into("<d", b, o, np.nextafter(pv, 1.0))
    assert pl.digest_mismatch(pl.stream_digest([bytes(b)] + streams[1:]), want) is None
    struct.pack_into("<d", b, o, pv * (1 + 1e-6))
    assert "p-values" in pl.digest_mismatch(pl.stream_digest([bytes(b)] + streams[1:]), want)
    b = bytearray(streams[0])
    b[4 + 33 + 9] ^= 1                           # a bit of the regression slope
    assert "other than a p-value" in pl.digest_mismatch(pl.stream_digest([bytes(b)] + streams[1:]), want)
