"""Target encoding on the device: `sample.ctdet.CenterDetectionSample`."""
