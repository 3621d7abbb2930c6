"""Reference-compatible import path: the (absent) reference keeps its networks in ``models/networks.py``
(pix2pixHD layout, README.md:101).  Everything here re-exports the B200 implementation."""
