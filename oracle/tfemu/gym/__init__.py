"""The gym symbols the reference imports (spaces.Box / Discrete, spaces.utils.flatdim) - TEST INFRASTRUCTURE, see
oracle/tfemu/README.md.  No simulators: ``make`` raises."""
from . import spaces  # noqa: F401


def make(name, **kw):
    raise ImportError("oracle/tfemu/gym has no simulators (asked for %r)" % (name,))
