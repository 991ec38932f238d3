"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's hot path, used as the checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
package (xnode-wan-pde-solver_b200/) never imports anything from here."""
