"""Payload adapters (reference utils/)."""
