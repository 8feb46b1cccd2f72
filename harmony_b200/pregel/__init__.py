"""Pregel (BSP graph processing) on the Elastic Table runtime."""
