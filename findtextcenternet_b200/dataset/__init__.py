"""train1 input pipeline on the device (mirror of the reference's ``dataset`` package for the detector: processer.pyx + data_detector.py)."""
