"""CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/oracle.h). Never imported by tracking_sdf_b200."""
