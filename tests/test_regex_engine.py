"""The dictionary-side regex matcher (csrc/lk_regex.cpp: RE2-syntax subset, partial match, 'i' flag) against Python's
re on the syntax both share.  regexp_matches(col, 're', 'i') -- BaseExpr.scala:485-486, 500-501."""
import ctypes
import re

import pytest

import helpers as H

PATTERNS = [
    "^pod-[0-4].*", "pod", "^$", "a|b|cd", "(ab)+c", "a{2,3}b", "a{2}", "a{2,}", "x*", "[^a-z]+", r"\d+\.\d+", r"\bfoo\b", r"\w+@\w+\.com",
    "^(GET|POST) /api/v[12]/", "colou?r", ".*compressed.*", r"\s", "[[:alpha:]]+[[:digit:]]", "(?i)ABC", "(?:ab|cd)ef$", "a.c", r"\x41B",
    r"\Qa.b\E", "[a-c-]+z", "é+", "^.{3}$", "(a|ab)(c|bcd)(d*)", "[]a]+", r"[\d\-]+", "a**", "((a)|(b))*c", "x{0}y", "^svc-0[37]$",
]
STRINGS = ["", "pod-3-abc", "POD-9", "pod-03", "aab", "aaab", "ab", "abab c", "ababc", "3.14", "foo", "a foo b", "afoob", "me@x.com", "GET /api/v2/x",
           "post /API/v1/", "color", "colour", "was compressed here", " ", "ab1", "abc", "ABC", "cdef", "abef", "a\nc", "AB", "a.b", "axb", "a-z", "éé",
           "日本語", "abcd", "]a]", "12-3", "c", "bac", "y", "xy", "svc-03", "svc-07", "svc-030"]


def _ours(p, s, ci=True):
    lib = H.emul_lib()
    lib.lk_emul_regex.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    b = s.encode()
    return lib.lk_emul_regex(p.encode(), 1 if ci else 0, b, len(b))


@pytest.mark.parametrize("pattern", PATTERNS)
def test_matches_python_re(pattern):
    py = pattern.replace("[[:alpha:]]", "[A-Za-z]").replace("[[:digit:]]", "[0-9]")
    py = {r"\Qa.b\E": r"a\.b", "a**": "a*"}.get(py, py)  # RE2-only spellings
    for s in STRINGS:
        got = _ours(pattern, s)
        assert got in (0, 1), f"{pattern!r} rejected"
        want = 1 if re.search(py, s, re.IGNORECASE) else 0
        assert got == want, f"{pattern!r} on {s!r}: ours {got}, python {want}"


def test_case_sensitive_mode_and_rejections():
    assert _ours("abc", "xABCx", ci=False) == 0
    assert _ours("abc", "xabcx", ci=False) == 1
    for bad in [r"(a)\1", "(?=a)b", "(?<!a)b", "a++", "[a", "(a", r"\p{Greek}", "*a"]:
        assert _ours(bad, "aa") < 0, bad  # LK_ERR_UNSUPPORTED, like RE2 would refuse it
