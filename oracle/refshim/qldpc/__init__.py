"""Stand-in for qldpc (import-time dependency of the reference's decoding.py only)."""
