"""Import-compatible stand-in for the reference package ``MoMA`` (B200 implementation in moma_b200)."""
