"""B200-native ensemble Kalman update behind the agarbuno/ces calibrate API."""
