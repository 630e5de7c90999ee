"""TEST INFRASTRUCTURE ONLY - see oracle/README.md.  Nothing under gp_rto_ei_b200/ imports this."""
