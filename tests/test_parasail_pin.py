"""Pins the restated tie rules against a REAL parasail wherever one is importable
(oracle/pin_parasail.py); skipped in containers without the wheel, where DESIGN.md says
"parity unpinned" for the alignment boundary."""
import pytest

from oracle import pin_parasail


def test_real_parasail_detection_never_returns_the_stub():
    mod = pin_parasail.real_parasail()
    assert mod is None or "parasail_stub" not in (getattr(mod, "__file__", "") or "")


def test_comparison_machinery_with_the_stub_standing_in():
    """The stub IS the committed rules: every other combination of the switches must be told
    apart from it by the case set (otherwise the cases could not pin anything)."""
    import sys
    sys.path.insert(0, str(pin_parasail._STUB))
    saved = sys.modules.pop("parasail", None)
    try:
        import parasail as stub
        report, committed, n = pin_parasail.compare(stub, pin_parasail.cases(n_random=80, max_len=70))
    finally:
        sys.path.remove(str(pin_parasail._STUB))
        sys.modules.pop("parasail", None)
        if saved is not None:
            sys.modules["parasail"] = saved
    assert n > 80 and report[committed] == 0
    assert all(bad > 0 for combo, bad in report.items() if combo != committed), report


def test_committed_rules_reproduce_real_parasail():
    parasail = pin_parasail.real_parasail()
    if parasail is None:
        pytest.skip("parasail is not installed here: the tie rules stay unpinned")
    report, committed, _ = pin_parasail.compare(parasail)
    assert report[committed] == 0, report
