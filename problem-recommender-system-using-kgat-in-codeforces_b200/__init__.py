"""kgat-b200 package (populated below)."""
