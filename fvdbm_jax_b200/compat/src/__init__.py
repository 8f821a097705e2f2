"""Stand-in for the reference's ``src`` package (see fvdbm_jax_b200.compat)."""
