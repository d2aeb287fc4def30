"""Sharding of the hot path across the GPUs of one box."""
