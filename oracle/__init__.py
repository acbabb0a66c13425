"""CPU oracle for the optik IK hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this package.  The product package
``optik_b200`` never does (tests/test_no_oracle_in_product.py enforces it).
"""
