"""Typed exceptions of the ET service (reference: services/et/.../exceptions/*.java)."""


class ETError(RuntimeError):
    """Base class of Elastic Table errors.  Reference: services/et/.../exceptions/*.java."""

    pass


class TableNotExistError(ETError):
    """Unknown table id.  Reference: services/et/.../exceptions/*.java."""

    pass


class BlockAlreadyExistsError(ETError):
    """putBlock of a block that is already stored.  Reference: services/et/.../exceptions/*.java."""

    pass


class BlockNotExistsError(ETError):
    """The block is not (or no longer) stored on this executor."""

    pass


class KeyGenerationError(ETError):
    """The local key space is exhausted or the table is not ordered.  Reference: services/et/.../exceptions/*.java."""

    pass


class NotAssociatedError(ETError):
    """The executor is not associated with the table.  Reference: services/et/.../exceptions/*.java."""

    pass


class AlreadyAssociatedError(ETError):
    """The executor is already associated with the table.  Reference: services/et/.../exceptions/*.java."""

    pass


class AlreadySubscribedError(ETError):
    """The executor already subscribes to the table.  Reference: services/et/.../exceptions/*.java."""

    pass


class NotSubscribedError(ETError):
    """The executor does not subscribe to the table.  Reference: services/et/.../exceptions/*.java."""

    pass


class ExecutorNotExistError(ETError):
    """Unknown executor id.  Reference: services/et/.../exceptions/*.java."""

    pass


class ChkpNotExistError(ETError):
    """Unknown checkpoint id / path.  Reference: services/et/.../exceptions/*.java."""

    pass


class PlanAlreadyExecutingError(ETError):
    """Only one plan runs at a time.  Reference: services/et/.../exceptions/*.java."""

    pass


class PlanOpExecutionError(ETError):
    """A plan op failed.  Reference: services/et/.../exceptions/*.java."""

    pass


class DataAccessFailedError(ETError):
    """A table access could not be completed.  Reference: services/et/.../exceptions/*.java."""

    pass


class TaskletFailedError(ETError):
    """A tasklet ended in the FAILED state.  Reference: services/et/.../exceptions/*.java."""

    pass
