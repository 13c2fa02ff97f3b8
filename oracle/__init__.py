"""CPU oracle -- test infrastructure only.  See the module headers for reference citations."""
