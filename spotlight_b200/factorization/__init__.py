"""Factorization models (reference: spotlight/factorization/)."""
