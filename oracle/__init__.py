"""CPU oracle for the ppo-radiotherapy environment-step path.

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.
"""
