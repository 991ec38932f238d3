"""TEST INFRASTRUCTURE ONLY -- provides `params['dim']` for /root/reference/configs/Ex4_3_funcs.py:3."""
params = {'dim': 5}
