/* TEST INFRASTRUCTURE: demos/pitch-tracking/pitch_detection.h includes <ffts/ffts.h> without using it */
