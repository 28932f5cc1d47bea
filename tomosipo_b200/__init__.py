"""tomosipo_b200 -- B200-native projector behind the tomosipo operator API."""
__version__ = "0.1.0"
