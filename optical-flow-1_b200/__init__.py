"""B200-native TV-L1 optical flow (hot path of 12334zq/optical-flow-1).  See DESIGN.md.

This directory is imported through the alias package `optical_flow_1_b200`.
"""
