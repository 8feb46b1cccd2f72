"""Long-running job server: accepts concurrent job submissions and schedules them on the pool."""
