"""Import-compatible stand-in for the reference package ``learning`` (B200 implementation in moma_b200)."""
