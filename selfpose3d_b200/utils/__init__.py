"""Host-side geometry helpers mirroring the reference's ``lib/utils/{cameras,transforms}.py``."""
