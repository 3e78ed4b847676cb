"""Command-line tools of the MCTF path (see cli.py)."""
