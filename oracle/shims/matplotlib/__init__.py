"""TEST INFRASTRUCTURE ONLY -- empty stand-in; the reference imports matplotlib.pyplot at
/root/reference/utils/auxillary_funcs.py:3 and uses it only inside `proj` (plotting, out of scope)."""
