"""sr_livo_b200 — B200-native LIO scan-matching hot path of SR-LIVO."""
