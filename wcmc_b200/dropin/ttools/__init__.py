"""Import shim: /root/reference/train_kpcn.py:34 imports ttools' crop_like but never calls it."""
