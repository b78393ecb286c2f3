"""tonic_b200: B200-native backend for Tonic's data-parallel hot path."""
