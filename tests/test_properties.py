"""Property tests (hypothesis) of the host-side codecs and partitioning helpers: encode -> decode round
trips and partition invariants on arbitrary inputs."""
import numpy as np
from hypothesis import given, settings, strategies as st

from remora_b200 import io, parallel
from remora_b200 import refine_signal_map as rsm

int16s = st.integers(min_value=-32768, max_value=32767)


@settings(max_examples=60, deadline=None)
@given(st.lists(int16s, min_size=0, max_size=600))
def test_vbz_round_trip_any_int16_sequence(values):
    sig = np.array(values, dtype=np.int16)
    assert np.array_equal(io.decode_vbz(io.encode_vbz(sig), sig.size), sig)


tag_values = st.one_of(
    st.tuples(st.just("i"), st.integers(-2 ** 31, 2 ** 31 - 1)),
    st.tuples(st.just("C"), st.integers(0, 255)),
    st.tuples(st.just("s"), st.integers(-2 ** 15, 2 ** 15 - 1)),
    st.tuples(st.just("Z"), st.text(alphabet="ACGTacgt0123456789:;,+?-_", max_size=30)),
    st.tuples(st.just("A"), st.sampled_from(list("PSI+-"))),
    st.tuples(st.just("Bc"), st.lists(st.integers(-128, 127), max_size=40)),
    st.tuples(st.just("BC"), st.lists(st.integers(0, 255), max_size=40)),
    st.tuples(st.just("Bs"), st.lists(st.integers(-2 ** 15, 2 ** 15 - 1), max_size=20)),
)
records = st.fixed_dictionaries({
    "query_name": st.text(alphabet="abcdef0123456789-", min_size=1, max_size=36),
    "query_sequence": st.text(alphabet="ACGTN", min_size=0, max_size=120),
    "flag": st.sampled_from([0, 4, 16, 256, 2048, 2064]),
    "mapping_quality": st.integers(0, 60),
    "tags": st.lists(tag_values, max_size=6),
})


@settings(max_examples=40, deadline=None)
@given(st.lists(records, min_size=1, max_size=8))
def test_bam_round_trip_any_records(tmp_path_factory, recs):
    path = str(tmp_path_factory.mktemp("bam") / "p.bam")
    out_recs = []
    for r in recs:
        tags = [(f"x{chr(97 + i)}", typ, (np.array(val) if typ[0] == "B" else val))
                for i, (typ, val) in enumerate(r["tags"])]
        n = len(r["query_sequence"])
        out_recs.append(dict(query_name=r["query_name"], query_sequence=r["query_sequence"], flag=r["flag"],
                             mapping_quality=r["mapping_quality"], reference_id=0, reference_start=7,
                             cigartuples=[(0, n)] if n else [], tags=tags))
    io.write_bam(path, "@HD\tVN:1.6\n@SQ\tSN:c\tLN:99999\n", [("c", 99999)], out_recs)
    with io.BamReader(path) as bam:
        got = list(bam)
    assert len(got) == len(out_recs)
    for g, w in zip(got, out_recs):
        assert g.query_name == w["query_name"] and g.query_sequence == w["query_sequence"]
        assert g.flag == w["flag"] and g.mapping_quality == w["mapping_quality"] and g.reference_start == 7
        assert g.cigartuples == w["cigartuples"] and [t for t, _ in g.tags] == [t for t, _, _ in w["tags"]]
        for (_, val), (_, typ, want) in zip(g.tags, w["tags"]):
            if typ[0] == "B":
                assert list(val) == list(want)
            else:
                assert val == want
        # SAM text -> fields survive too
        fields = g.to_sam().split("\t")
        assert fields[0] == w["query_name"] and fields[9] == (w["query_sequence"] or "*")


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 10 ** 6), max_size=200), st.integers(1, 16))
def test_shard_by_work_partitions(work, world):
    parts = [parallel.shard_by_work(work, world, r) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(len(work)))
    if work:
        loads = [sum(work[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(work)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 1000), max_size=100), st.integers(1, 5000))
def test_split_by_budget_partitions(sizes, budget):
    spans = rsm.split_by_budget(sizes, budget)
    assert [i for s, e in spans for i in range(s, e)] == list(range(len(sizes)))
    for s, e in spans:
        assert e > s and (sum(sizes[s:e]) <= budget or e - s == 1)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, width=32), min_size=1, max_size=300),
       st.lists(st.floats(min_value=0.0, max_value=1.0), min_size=1, max_size=8), st.booleans())
def test_quantile_linear_equals_numpy(values, qs, as_f32):
    a = np.array(values, dtype=np.float32 if as_f32 else np.float64)
    q = np.array(qs, dtype=np.float64)
    want, got = np.quantile(a, q), rsm.quantile_linear(a, q)
    assert want.dtype == got.dtype and np.array_equal(want, got)
