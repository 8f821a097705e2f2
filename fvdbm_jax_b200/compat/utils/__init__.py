"""Stand-in for the reference's ``utils`` package (see fvdbm_jax_b200.compat)."""
