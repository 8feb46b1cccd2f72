"""Typed exceptions of the ET service (reference: services/et/.../exceptions/*.java)."""


class ETError(RuntimeError):
    pass


class TableNotExistError(ETError):
    pass


class BlockAlreadyExistsError(ETError):
    pass


class BlockNotExistsError(ETError):
    pass


class KeyGenerationError(ETError):
    pass


class NotAssociatedError(ETError):
    pass


class AlreadyAssociatedError(ETError):
    pass


class AlreadySubscribedError(ETError):
    pass


class NotSubscribedError(ETError):
    pass


class ExecutorNotExistError(ETError):
    pass


class ChkpNotExistError(ETError):
    pass


class PlanAlreadyExecutingError(ETError):
    pass


class PlanOpExecutionError(ETError):
    pass


class DataAccessFailedError(ETError):
    pass


class TaskletFailedError(ETError):
    pass
