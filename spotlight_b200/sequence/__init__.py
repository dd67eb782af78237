"""Sequence models (reference: spotlight/sequence/)."""
