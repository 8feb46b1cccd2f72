"""ML applications on Dolphin: MLR, NMF, LDA, Lasso, GBT + the AddInteger / AddVector examples."""
