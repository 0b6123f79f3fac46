"""TEST INFRASTRUCTURE: CPU oracle (restatement of the reference's algorithm for the hot path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from here - as the checker, never as the thing shipped.
"""
