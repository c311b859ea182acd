"""Import stub for the reference's plotting imports (utils/*.py:7)."""
