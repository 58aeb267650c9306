"""Checker libraries (CPU restatement of the reference hot paths). Test infrastructure only."""
