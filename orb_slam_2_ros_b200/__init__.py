"""orb_slam_2_ros_b200 — B200-native (sm_100a) ORB front-end of ORB-SLAM2 behind the reference's own
interfaces: ORBextractor, ORBmatcher, ComputeStereoMatches, ORBVocabulary::transform.  See DESIGN.md / INTEGRATION.md."""
from .extractor import ORBextractor  # noqa: F401
from .matcher import DescriptorDB, ORBmatcher, hamming_top2, hamming_top2_csr, top2_merge, top2_merge_device  # noqa: F401
from .stereo import compute_stereo_matches  # noqa: F401
from .vocabulary import ORBVocabulary  # noqa: F401
from . import map_records  # noqa: F401
