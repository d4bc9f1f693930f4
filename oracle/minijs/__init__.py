"""TEST INFRASTRUCTURE ONLY.  A small ECMAScript-subset interpreter, written for one purpose: to EXECUTE THE
REFERENCE'S OWN minified modules (dist/main.js, webpack module 584 = formantanalyzer@1.1.6, inner modules 0/3/4/7)
in an image that has no JS engine, so that the C oracle can be pinned against outputs of the reference's code
rather than against a second restatement.  See oracle/minijs/run_reference.py."""
from .interp import Interp, JSArray, JSObject, JSTyped, JSThrow, UNDEF  # noqa: F401
