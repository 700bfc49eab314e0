"""Test infrastructure: CPU oracle for the HICom compressor path.  Never imported by ``hicom_b200``."""
