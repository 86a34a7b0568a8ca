"""Oracle restatement of the per-segment stream element wire format: producer
Commons.scala:474-502 + SSEMessage.scala:23-34, consumer SegmentSequencer.scala:35-101."""
import json
import math

import pytest

import lakeside_oracle as lo


def test_sse_shape_matches_reference_envelope():
    b = lo.to_sse([lo.SketchInput(1700000000000, {"resource.service.name": "svc-03"}, {"sum": 2.5})])
    assert b.endswith(b"\r\n\r\n") and b.startswith(b"data: ")
    obj = json.loads(b[len(b"data: "):].decode())
    # GenericSSEPayload(id = "_", type = "data", message = Map(timestamp, tags, type, sketchType, sketch))
    assert obj["id"] == "_" and obj["type"] == "data"
    assert obj["message"] == {"timestamp": 1700000000000, "tags": {"resource.service.name": "svc-03"}, "type": "sketch",
                              "sketchType": "map", "sketch": {"sum": 2.5}}


def test_sse_round_trip_with_non_finite_and_escapes():
    els = [
        lo.SketchInput(10, {"a": 'q"uo\\te', "b": "line\nbreak\ttab", "c": "ünï"}, {"sum": 1.0 / 3.0, "min": math.nan, "max": math.inf}),
        lo.SketchInput(10, {}, {"count": -math.inf}),
        lo.DataPoint(timestamp=11, value=5e-324, tags={"k": "v"}),
    ]
    got = lo.sse_decode(lo.to_sse(els))
    assert len(got) == 3
    assert got[0].tags == els[0].tags and got[0].timestamp == 10
    assert got[0].sketch["sum"] == 1.0 / 3.0 and math.isnan(got[0].sketch["min"]) and got[0].sketch["max"] == math.inf
    assert got[1].sketch == {"count": -math.inf} and got[1].tags == {}
    assert isinstance(got[2], lo.DataPoint) and got[2].value == 5e-324 and got[2].tags == {"k": "v"}


def test_decoder_tolerances():
    # SegmentSequencer.scala:35-51: numbers as strings, "nan", "+Infinity"; unparsable -> NaN / 0
    ev = ('data: {"id":"_","type":"data","message":{"timestamp":"42","tags":{},"type":"sketch","sketchType":"map",'
          '"sketch":{"a":"nan","b":"+Infinity","c":"1.5","d":"zzz","e":null}}}\r\n\r\n'
          'data: {"type":"heartbeat"}\r\n\r\n')
    (s,) = lo.sse_decode(ev.encode())
    assert s.timestamp == 42
    assert math.isnan(s.sketch["a"]) and s.sketch["b"] == math.inf and s.sketch["c"] == 1.5
    assert math.isnan(s.sketch["d"]) and math.isnan(s.sketch["e"])
    bad = 'data: {"id":"_","type":"data","message":{"timestamp":1,"tags":{"k":5},"type":"sketch","sketchType":"map","sketch":{}}}\r\n\r\n'
    with pytest.raises(lo.OracleQueryError):
        lo.sse_decode(bad.encode())
