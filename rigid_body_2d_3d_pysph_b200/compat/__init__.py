"""Host-side mirror of the PySPH surface the reference scripts import."""
