"""Mirror of the reference lib/ package for the training-loop hot path."""
