"""Stub of `ultralytics` (absent from the image; un-pinned dependency of the
reference).  Only the four names the reference imports exist.  TEST INFRASTRUCTURE."""
