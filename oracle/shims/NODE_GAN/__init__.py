"""TEST INFRASTRUCTURE ONLY -- /root/reference/configs/Ex4_3_funcs.py:3 imports `NODE_GAN.main.params`."""
