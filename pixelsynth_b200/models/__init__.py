"""Host-side mirror of the reference package `models` (crockwell/pixelsynth), hot path only."""
