"""TEST INFRASTRUCTURE — CPU restatement of the reference SPF (see oracle/*.cc).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (holo_b200) never does.
"""
