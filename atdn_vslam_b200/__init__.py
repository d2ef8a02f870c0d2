"""atdn_vslam_b200 -- B200-native (sm_100a) odometry front end of ATDN vSLAM.

Drop-in classes keep the reference's Python signatures:
``RAFTGMA``, ``CorrBlock`` (GMA flow), ``ATDNVO`` (CLVO pose), ``KeyframeIndex``/``MappingEncoder``
(localization).  All compute goes through ``libatdn_b200.so`` (C ABI, ``include/atdn_b200.h``);
there is no CPU or PyTorch fallback -- a missing library or non-sm_100 device raises.
"""
__version__ = "0.1.0"
