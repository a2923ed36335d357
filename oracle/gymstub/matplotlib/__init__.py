"""Empty matplotlib stand-in: the reference imports it at module scope for render() only."""
