"""Drop-in mirror of the reference's `libdl` package surface for the hot path."""
