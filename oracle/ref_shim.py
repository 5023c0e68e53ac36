"""
TEST INFRASTRUCTURE — makes the UNMODIFIED reference (/root/reference, AtomAI v0.8.1) importable in
the build container, where matplotlib / skimage / mendeleev / gpytorch are not installed
(SURVEY.md §8c, Appendix A).  Used only by tests/golden/make_golden*.py to generate the committed
golden vectors; nothing under atomai_b200/ imports it and it never runs on the GPU box
(/root/reference does not exist there).  The stub set lives in baseline/ref_loader.py, which the
bench's reference arms share (they load the vendored copy under baseline/_ref instead).
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def import_reference():
    """Returns the reference `atomai` package (CPU), imported from /root/reference."""
    from baseline.ref_loader import import_reference as _imp
    return _imp(REFERENCE_ROOT)
