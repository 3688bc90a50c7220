"""Mirror of the reference ``src/utility`` package (only ``utils`` is on the hot path)."""
