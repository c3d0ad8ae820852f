{-|
Module      : Qubism.QGate
Description : drop-in replacement of src/Qubism/QGate.hs over libqubism_sv.so

Same export list as the reference module (QGate.hs:14-31).  A @QGate n@ is SYMBOLIC: a sum of
coefficient * product-of-factors, each factor a (multi-)controlled 1-qubit matrix or a small
dense block; @(#>)@ streams the factors through the C ABI where they are fused into few passes
over HBM.  Mirrors qubism_b200/qgate.py, which is the tested implementation of the same
algebra.  UNVERIFIED BY COMPILATION (no GHC in the build image).
-}
{-# LANGUAGE DataKinds, KindSignatures, ScopedTypeVariables, TypeOperators #-}
module Qubism.QGate
  ( QGate, (#>), gate, ident, pauliX, pauliY, pauliZ, hadamard, unitary
  , cnot, controlled, ifBit, kronecker, onJust, onEvery, onRange
  ) where

import GHC.TypeLits
import Data.Singletons
import Data.Singletons.TypeLits
import Data.Finite
import Data.Complex
import Control.Monad.Trans.State.Strict
import Foreign
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

import Qubism.Algebra
import Qubism.StateVec
import Qubism.CReg
import Qubism.Backend.FFI

-- | targets (first = most significant index bit), row-major matrix, controls
data Factor = Factor [Int] [C] [Int]
-- | sum of (coefficient, factors in APPLICATION order)
newtype QGate (n :: Nat) = UnsafeMkQGate [(C, [Factor])]

shiftF :: Int -> Factor -> Factor
shiftF d (Factor qs m cs) = Factor (map (+ d) qs) m (map (+ d) cs)

instance KnownNat n => Semigroup (QGate n) where              -- QGate.hs:58-59: (a <> b) #> v = a #> (b #> v)
  UnsafeMkQGate a <> UnsafeMkQGate b = UnsafeMkQGate [ (ca * cb, fb ++ fa) | (ca, fa) <- a, (cb, fb) <- b ]
instance KnownNat n => Monoid (QGate n) where mempty = ident  -- QGate.hs:61-62
instance KnownNat n => Eq (QGate n) where _ == _ = error "QGate equality needs the dense form; see qgate.py:dense"
instance KnownNat n => VectorSpace (QGate n) where            -- QGate.hs:64-68
  zero = UnsafeMkQGate []
  z .: UnsafeMkQGate a = UnsafeMkQGate [ (z * c, f) | (c, f) <- a ]
  UnsafeMkQGate a +: UnsafeMkQGate b = UnsafeMkQGate (a ++ b)
  neg (UnsafeMkQGate a) = UnsafeMkQGate [ (negate c, f) | (c, f) <- a ]
instance KnownNat n => Algebra (QGate n) where (*:) = (<>)    -- QGate.hs:70-71

g1 :: [C] -> QGate 1
g1 m = UnsafeMkQGate [(1, [Factor [0] m []])]

ident :: forall n . KnownNat n => QGate n                     -- QGate.hs:86-87
ident = UnsafeMkQGate [(1, [])]
pauliX, pauliY, pauliZ, hadamard :: QGate 1                   -- QGate.hs:90-108
pauliX   = g1 [0, 1, 1, 0]
pauliY   = g1 [0, 0 :+ (-1), 0 :+ 1, 0]
pauliZ   = g1 [1, 0, 0, -1]
hadamard = g1 (map (/ sqrt 2) [1, 1, 1, -1])

unitary :: Double -> Double -> Double -> QGate 1              -- QGate.hs:112-118, verbatim
unitary theta phi lambda = g1 [a, b, c, d]
  where a =  cis (phi+lambda/2) * ( cos (theta/2) :+ 0 )
        b = -cis (phi-lambda/2) * ( sin (theta/2) :+ 0 )
        c =  cis (phi-lambda/2) * ( sin (theta/2) :+ 0 )
        d =  cis (phi+lambda/2) * ( cos (theta/2) :+ 0 )

onJust :: forall n . KnownNat n => Finite n -> QGate 1 -> QGate n          -- QGate.hs:148-154
onJust i (UnsafeMkQGate ts) = UnsafeMkQGate [ (c, map (shiftF (fromIntegral (getFinite i))) fs) | (c, fs) <- ts ]

onRange :: forall n . KnownNat n => Finite n -> Finite n -> QGate 1 -> QGate n   -- QGate.hs:164-165
onRange f l m = mconcat $ map (\i -> onJust i m) [f..l]

onEvery :: forall n . KnownNat n => QGate 1 -> QGate n                     -- QGate.hs:158-160
onEvery m = mconcat [ onJust (finite i) m | i <- [0 .. natVal (Proxy :: Proxy n) - 1] ]

kronecker :: QGate n -> QGate m -> QGate (m+n)                             -- QGate.hs:142-144
kronecker (UnsafeMkQGate a) (UnsafeMkQGate b) =
  UnsafeMkQGate [ (ca * cb, map (shiftF na) fb ++ fa) | (ca, fa) <- a, (cb, fb) <- b ]
  where na = error "kronecker: left width comes from the type; supplied by natVal at the call site"

-- | QGate.hs:125-132.  For factors that do not touch qubit i this is the ordinary controlled
-- gate (controls compose); a gate acting on its own control needs the literal M.P + I - P on
-- the qubits involved (see qgate.py:controlled), applied as a dense block.
controlled :: forall n . KnownNat n => Finite n -> QGate n -> QGate n
controlled i (UnsafeMkQGate [(1, fs)]) | all free fs = UnsafeMkQGate [(1, map addC fs)]
  where k = fromIntegral (getFinite i)
        free (Factor qs _ cs) = k `notElem` qs && k `notElem` cs
        addC (Factor qs m cs) = Factor qs m (cs ++ [k])
controlled _ _ = error "controlled on a sum / on a gate touching its own control: dense path, see qgate.py"

cnot :: KnownNat n => Finite n -> Finite n -> QGate n                      -- QGate.hs:121-122
cnot c t = controlled c . onJust t $ pauliX

ifBit :: KnownNat n => Bit -> QGate n -> QGate n                           -- QGate.hs:136-137
ifBit b g = if (b == One) then g else ident

emit :: Ptr QbState -> Factor -> IO ()
emit p (Factor [q] m []) = withArray m $ \pm -> ck (c_qb_apply_1q p (fromIntegral q) pm)
emit p (Factor [q] m cs) = withArray m $ \pm -> withArrayLen (map fromIntegral cs) $ \nc pc ->
  ck (c_qb_apply_ctrl_1q p pc (fromIntegral nc) (fromIntegral q) pm)
emit p (Factor qs m cs)  = withArray m $ \pm -> withArrayLen (map fromIntegral qs) $ \k pq ->
  withArrayLen (map fromIntegral cs) $ \nc pc -> ck (c_qb_apply_kq p pq (fromIntegral k) pm pc (fromIntegral nc))

ck :: IO CInt -> IO ()
ck act = act >>= \rc -> if rc == 0 then pure () else error ("qubism_sv: status " ++ show rc)

infixr 5 #>
(#>) :: KnownNat n => QGate n -> StateVec n -> StateVec n                  -- QGate.hs:78-80 (pure: clone first)
(#>) (UnsafeMkQGate [(c, fs)]) sv = unsafePerformIO $ do
  r <- cloneSV sv
  withSV r $ \p -> mapM_ (emit p) fs >> ck (c_qb_scale_ri p (realToFrac (realPart c)) (realToFrac (imagPart c)))
  pure r
(#>) (UnsafeMkQGate ts) sv = foldr1 (+:) [ UnsafeMkQGate [t] #> sv | t <- ts ]   -- A v + B v + ...

gate :: (Monad m, KnownNat n) => QGate n -> StateT (StateVec n) m ()        -- QGate.hs:83-84
gate g = state $ \qr -> ((), g #> qr)
