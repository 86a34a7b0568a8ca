"""GPU: the CUDA path through the C ABI (lk_eval) on every edge case of tests/cases.py, against the oracle."""
import pytest

import cases
import helpers as H

pytestmark = pytest.mark.gpu
CASES = cases.all_cases() + cases.gpu_only_cases()
ERRORS = cases.error_cases()


@pytest.fixture(scope="module", autouse=True)
def _init():
    from lakeside_b200 import api

    api.init()


@pytest.mark.parametrize("cid,paths,rq,ops", CASES, ids=[c[0] for c in CASES])
def test_gpu_matches_oracle(cid, paths, rq, ops):
    want = H.oracle_single(rq, paths)
    got = H.gpu_eval_single(rq, paths)
    H.assert_same(got, want, ops, cid)
    assert got["cols"][0] == want["cols"][0] and got["cols"][2:] == want["cols"][2:]


@pytest.mark.parametrize("cid,paths,rq,ops", CASES[::3], ids=[c[0] for c in CASES[::3]])
def test_gpu_hash_path_matches_oracle(cid, paths, rq, ops):
    from lakeside_b200 import api

    want = H.oracle_single(rq, paths)
    with api.Query(rq, path="hash") as q:
        for p in paths:
            q.add_segment_bytes(open(p, "rb").read())
        q.prepare()
        q.execute()
        res = q.finalize()
        got = H.canon_from_gpu(res)
        res.close()
    H.assert_same(got, want, ops, cid + " (hash)")


@pytest.mark.parametrize("cid,paths,rq,kind", ERRORS, ids=[c[0] for c in ERRORS])
def test_gpu_error_codes(cid, paths, rq, kind):
    from lakeside_b200 import api

    code = {"invalid": 1, "unsupported": 2, "query": 5}[kind]
    with pytest.raises(api.LakesideError) as e:
        api.eval_glob(rq, paths)
    assert e.value.code == code, str(e.value)


def test_jdbc_style_row_access_and_data_points():
    # what Commons.toDataPoint does with the ResultSet (Commons.scala:399-462)
    from lakeside_b200 import api
    import lakeside_oracle as lo

    cid, paths, rq, ops = next(c for c in CASES if c[0] == "ops/agg_count")
    res = api.eval_glob(rq, paths)
    assert res.columns[0] == "step_ts" and res.columns[2] == "name"
    want = lo.evaluate_glob(lo.push_down_request_from_json(rq), paths)
    dps = res.to_data_points({"q": "tags"})
    odps = lo.to_data_points(want, {"q": "tags"})
    key = lambda d: (d.timestamp, tuple(sorted(d.tags.items())))
    assert sorted(map(key, dps)) == sorted(map(key, odps))
    # dropped tags ("" / "null" / NULL) can make several rows share one tag map: compare as multisets
    assert sorted((key(d), d.value) for d in dps) == sorted((key(d), d.value) for d in odps)
    assert all("missing.group" not in d.tags for d in dps)
    for row in range(min(res.num_rows, 50)):
        assert res.get_long(row, 1) == int(res.ts[row])
        assert res.get_double(row, 2) == float(res.values[0][row])
        for t in range(res.num_tags):
            c = res.tag_codes[t][row]
            assert res.get_string(row, 3 + t) == (None if c < 0 else res.tag_dicts[t][c])
    res.close()
