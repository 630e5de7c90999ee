"""Host-side stand-ins for third-party modules the reference imports but this image lacks."""
