"""Shared services: parameters, centralized communication, data loading/storing, metrics."""
