"""Minimal stand-ins for the THIRD-PARTY packages the reference's pixel decoder imports
(msdeformattn.py:10,17-19).  Test infrastructure only; not reference code and not product code."""
