"""srb200 - host side of the B200-native incremental-session path (ctypes over libsrb200.so)."""
