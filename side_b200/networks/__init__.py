from .stereo_network import cost_volume, get_pose_net, get_proposal_shift, stereo_network  # noqa: F401
from .feature_extraction_dla34 import DeformConv, feature_extraction_dla34  # noqa: F401
