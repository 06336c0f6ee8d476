"""Loops, losses, decode and configuration of the mirror (reference lib/core)."""
