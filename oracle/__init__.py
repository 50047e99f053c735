"""oracle/ — CPU restatements of the reference's algorithm.  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs, never by the
product package (tce-rvos_b200/)."""
