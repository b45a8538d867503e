"""Host-side commons mirroring ``chattts_plus/commons`` (reference) for the hot path's callers."""
