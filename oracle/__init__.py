"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot-path arithmetic, used solely as the
parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under stable-diffusion-on-device_b200/ imports
this package; the product path fails loudly without its CUDA library instead of
falling back here.
"""
