"""B200-native training-loop hot path of EpipolarPose: C ABI (libepb.so), engines, and the reference-shaped Python surface."""
