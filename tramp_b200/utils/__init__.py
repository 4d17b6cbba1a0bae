"""Numerical helpers of the EP path that the reference keeps under tramp/utils/: the
truncated-normal moments (truncated_normal.py), evaluated on the device."""
