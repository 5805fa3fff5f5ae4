"""CPU oracles -- test infrastructure only.  See the header of each module.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import from here.  The product package never does.
"""
