"""Stub: the reference imports matplotlib at module top but the hot path never plots."""
