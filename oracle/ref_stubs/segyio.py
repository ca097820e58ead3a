"""empty stand-in: SEG-Y ingest is not on the hot path"""
