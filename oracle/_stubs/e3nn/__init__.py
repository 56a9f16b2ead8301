"""Stub of e3nn (absent offline; TEST INFRASTRUCTURE).  It exists only so that the reference's
src/misc/sh_rotation.py can be IMPORTED by oracle/epipolar_ref.py; calling into it raises."""
