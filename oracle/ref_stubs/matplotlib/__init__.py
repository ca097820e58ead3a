"""empty stand-in: plotting is not on the hot path"""
cm = None
