"""``from models.networks import define_G, define_D`` — the reference's factory API (BASELINE.json names
it; pix2pixHD signatures), served by the sm_100a implementation in nhvr_b200.networks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nhvr_b200.networks import define_G, define_D, GlobalGeneratorB200, MultiscaleDiscriminatorB200  # noqa: F401,E402
